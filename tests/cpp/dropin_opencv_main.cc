// TEST PROGRAM: the drop-in classes in their OpenCV mode (BRISK_B200_USE_OPENCV), held and called through the cv::
// base classes the way an OpenCV 3 application does.  OpenCV itself is not installed in this image: the test compiles
// against the oracle's minimal stand-in headers (oracle/shim, test infrastructure), which declare the same cv::Feature2D /
// cv::DescriptorMatcher virtuals the reference overrides.
//   usage: dropin_opencv_main <image.pgm> <out.bin>
#define BRISK_B200_USE_OPENCV 1
#include <brisk/brisk.h>

#include <cstdio>
#include <fstream>
#include <iostream>

static cv::Mat ReadPgm(const char* path) {
  std::ifstream f(path, std::ios::binary);
  std::string magic; int w, h, maxv;
  f >> magic >> w >> h >> maxv;
  f.get();
  cv::Mat m(h, w, CV_8UC1);
  f.read(reinterpret_cast<char*>(m.data), (std::streamsize)w * h);
  return m;
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  cv::Mat img = ReadPgm(argv[1]);
  // held through the base classes, like any other cv::Feature2D / cv::DescriptorMatcher
  cv::Ptr<cv::Feature2D> detector(new cv::BriskFeatureDetector(70));          // alias of brisk/brisk.h:56-59
  cv::Ptr<cv::Feature2D> extractor(new cv::BriskDescriptorExtractor());
  cv::Ptr<cv::Feature2D> harris(new brisk::HarrisScaleSpaceFeatureDetector(0, 30.0, 20.0));
  cv::Ptr<cv::Feature2D> both(new brisk::BriskFeature(0, 30.0, 20.0));
  cv::Ptr<cv::DescriptorMatcher> matcher(new brisk::BruteForceMatcherSse());

  std::vector<cv::KeyPoint> kps, hk, bk;
  cv::Mat desc, bdesc;
  detector->detect(img, kps);                 // cv::Feature2D::detect -> detectAndCompute override
  extractor->compute(img, kps, desc);         // cv::Feature2D::compute -> detectAndCompute(useProvidedKeypoints)
  harris->detect(img, hk);
  both->detectAndCompute(img, cv::noArray(), bk, bdesc);

  matcher->add(std::vector<cv::Mat>(1, desc));
  std::vector<std::vector<cv::DMatch> > knn;
  matcher->knnMatch(desc, knn, 2);            // cv::DescriptorMatcher::knnMatch -> knnMatchImpl override
  int self = 0;
  for (size_t i = 0; i < knn.size(); ++i) self += knn[i].size() == 2 && knn[i][0].distance == 0.0f && knn[i][0].trainIdx <= (int)i;  // (a duplicate row matches its first copy)
  cv::Ptr<cv::DescriptorMatcher> copy = matcher->clone();
  std::vector<std::vector<cv::DMatch> > rad;
  copy->radiusMatch(desc, rad, 40.0f);
  long long nrad = 0;
  for (const auto& v : rad) nrad += (long long)v.size();

  // HarrisScoreCalculator's public accessors
  brisk::HarrisScoreCalculator calc;
  calc.SetImage(img);
  std::vector<brisk::HarrisScoreCalculator::PointWithScore> maxima;
  calc.Get2dMaxima(maxima, 20);

  FILE* f = std::fopen(argv[2], "wb");
  const int32_t head[8] = {(int32_t)kps.size(), desc.cols, self, (int32_t)hk.size(), (int32_t)bk.size(), (int32_t)nrad, (int32_t)maxima.size(),
                           calc.Score(100, 120)};
  std::fwrite(head, 4, 8, f);
  const double sd = calc.Score(100.25, 120.5);
  std::fwrite(&sd, 8, 1, f);
  std::fwrite(kps.data(), sizeof(cv::KeyPoint), kps.size(), f);
  std::fwrite(desc.data, 1, (size_t)desc.rows * desc.cols, f);
  std::fwrite(hk.data(), sizeof(cv::KeyPoint), hk.size(), f);
  std::fwrite(bdesc.data, 1, (size_t)bdesc.rows * bdesc.cols, f);
  for (const auto& m : maxima) { const int32_t t[3] = {m.score, m.x, m.y}; std::fwrite(t, 4, 3, f); }
  std::fclose(f);
  return 0;
}
