// Exercises the header-only C++ drop-in (include/brisk/brisk.h) the way an OKVIS-style caller
// uses the reference: detector + extractor per image, then a brute-force match.
// usage: dropin_main <in.pgm> <out.bin> [<matches.bin>]
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>

#include <brisk/brisk.h>

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  std::ifstream f(argv[1], std::ios::binary);
  std::string magic; int w, h, maxv;
  f >> magic >> w >> h >> maxv;
  f.get();
  agast::Mat img(h, w, CV_8UC1);
  f.read(reinterpret_cast<char*>(img.data), (std::streamsize)w * h);
  brisk::BriskFeatureDetector detector(70);
  brisk::BriskDescriptorExtractor extractor;
  std::vector<agast::KeyPoint> kps;
  detector.detect(img, kps);
  agast::Mat desc;
  extractor.compute(img, kps, desc);
  brisk::BruteForceMatcher matcher;
  std::vector<std::vector<brisk::DMatch> > matches;
  matcher.knnMatch(desc, desc, matches, 2);
  int self = 0;
  for (size_t i = 0; i < matches.size(); ++i) self += (matches[i].size() == 2 && matches[i][0].distance == 0);
  // Harris scale-space detector + extractor in one object (reference test-binary-equal.cc:73-89)
  brisk::BriskFeature feature(0, 30.0, 20.0);
  std::vector<agast::KeyPoint> hk;
  agast::Mat hd;
  feature.detectAndCompute(img, agast::Mat(), hk, hd);
  std::printf("harris: %d key points\n", (int)hk.size());
  std::ofstream o(argv[2], std::ios::binary);
  int n = (int)kps.size(), nb = desc.cols;
  o.write(reinterpret_cast<char*>(&n), 4); o.write(reinterpret_cast<char*>(&nb), 4); o.write(reinterpret_cast<char*>(&self), 4);
  o.write(reinterpret_cast<char*>(kps.data()), (std::streamsize)n * sizeof(agast::KeyPoint));
  o.write(reinterpret_cast<char*>(desc.data), (std::streamsize)n * nb);
  int hn = (int)hk.size();
  o.write(reinterpret_cast<char*>(&hn), 4);
  o.write(reinterpret_cast<char*>(hk.data()), (std::streamsize)hn * sizeof(agast::KeyPoint));
  o.write(reinterpret_cast<char*>(hd.data), (std::streamsize)hn * hd.cols);
  // ComputeScale: the detected key points re-examined in every layer (reference brisk-feature-detector.cc:87-92)
  std::vector<agast::KeyPoint> cs(kps);
  detector.ComputeScale(img, cs);
  int cn = (int)cs.size();
  o.write(reinterpret_cast<char*>(&cn), 4);
  o.write(reinterpret_cast<char*>(cs.data()), (std::streamsize)cn * sizeof(agast::KeyPoint));
  // Harris detector on a non-empty vector: "use passed key points" (reference scale-space-feature-detector.h:103-108)
  brisk::HarrisScaleSpaceFeatureDetector hdet(0, 30.0, 20.0);
  std::vector<agast::KeyPoint> pk(hk);
  hdet.detect(img, pk);
  int pn = (int)pk.size();
  o.write(reinterpret_cast<char*>(&pn), 4);
  o.write(reinterpret_cast<char*>(pk.data()), (std::streamsize)pn * sizeof(agast::KeyPoint));
  // the distance primitive's static form (hamming.h:79-91) and the detectAndCompute overrides
  int ham = n > 1 ? (int)brisk::Hamming::PopcntofXORed(desc.data, desc.data + desc.step, nb / 16) : -1;
  o.write(reinterpret_cast<char*>(&ham), 4);
  std::vector<agast::KeyPoint> k2;
  agast::Mat none, d2;
  detector.detectAndCompute(img, agast::Mat(), k2, none);
  extractor.detectAndCompute(img, agast::Mat(), k2, d2);
  int same = k2.size() == kps.size() && d2.rows == desc.rows && std::memcmp(d2.data, desc.data, (size_t)desc.rows * nb) == 0;
  o.write(reinterpret_cast<char*>(&same), 4);
  std::printf("%d key points, %d-byte descriptors, %d self matches, %d from ComputeScale, %d of the passed Harris points\n", n, nb, self, cn, pn);
  if (argc > 3) {
    // matcher surface: a two-image train collection with masks, knnMatch(k = 3) and radiusMatch(45)
    const int nq = std::min(n, 150), n0 = n / 3;
    agast::Mat q(nq, nb, CV_8UC1, desc.data, desc.step);
    std::vector<agast::Mat> train;
    train.push_back(agast::Mat(n0, nb, CV_8UC1, desc.data, desc.step));
    train.push_back(agast::Mat(n - n0, nb, CV_8UC1, desc.data + (size_t)n0 * desc.step, desc.step));
    std::vector<agast::Mat> masks;
    masks.push_back(agast::Mat::zeros(nq, n0, CV_8UC1));
    masks.push_back(agast::Mat::zeros(nq, n - n0, CV_8UC1));
    for (int i = 0; i < nq; ++i) {
      for (int t = 0; t < n0; ++t) masks[0].data[(size_t)i * n0 + t] = (i % 7 != 3) && ((i + t) % 3 != 0);
      for (int t = 0; t < n - n0; ++t) masks[1].data[(size_t)i * (n - n0) + t] = (i % 7 != 3) && ((i + 2 * t) % 5 != 0);
    }
    brisk::BruteForceMatcher coll;
    coll.add(train);
    std::ofstream mo(argv[3], std::ios::binary);
    auto dump = [&](const std::vector<std::vector<brisk::DMatch> >& mm) {
      int lists = (int)mm.size();
      mo.write(reinterpret_cast<char*>(&lists), 4);
      for (const auto& v : mm) {
        int c = (int)v.size();
        mo.write(reinterpret_cast<char*>(&c), 4);
        for (const auto& m : v) mo.write(reinterpret_cast<const char*>(&m), sizeof(brisk::DMatch));
      }
    };
    std::vector<std::vector<brisk::DMatch> > mm;
    coll.knnMatch(q, mm, 3, masks, false);
    dump(mm);
    coll.radiusMatch(q, mm, 45.0f, masks, true);
    dump(mm);
  }
  return 0;
}
