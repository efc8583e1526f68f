// ORACLE SCAFFOLDING (test infrastructure, never linked into the product).
//
// C entry points around the UNMODIFIED reference sources (compiled in place
// from /root/reference by oracle/Makefile into oracle/_ref/libbrisk_ref.so).
// Used (a) to pin the CPU restatement in oracle/brisk_oracle.cc, (b) as the
// parity checker of the CUDA path in tests/, and (c) as the "reference" CPU
// baseline that bench.py times on the GPU box's host cores.
//
// Private members of the reference classes are reached with the
// `#define private public` trick so that stage outputs (pyramid layers,
// threshold map, raw AGAST corners, pattern look-up table) can be dumped.

#include <stdint.h>
#include <algorithm>
#include <bitset>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <istream>
#include <limits>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include <opencv2/core/core.hpp>
#include <opencv2/features2d/features2d.hpp>
#include <agast/glog.h>

#define private public
#define protected public
#include <brisk/internal/brisk-layer.h>
#include <brisk/internal/brisk-scale-space.h>
#include <brisk/brisk-descriptor-extractor.h>
#include <brisk/brisk-feature-detector.h>
#include <brisk/harris-score-calculator.h>
#include <brisk/scale-space-feature-detector.h>
#include <brisk/brute-force-matcher.h>
#include <brisk/harris-feature-detector.h>
// The legacy detector's stage functions are `__inline__` members defined in its .cc: the unmodified source file is
// compiled as part of this translation unit (instead of on its own) so that the stage dump below can call them.
#include <../src/harris-feature-detector.cc>
#undef private
#undef protected
#include <brisk/internal/harris-scores.h>
#include <brisk/internal/hamming.h>
#include <brisk/internal/image-down-sampling.h>
// integral-image.h defines its functions non-inline (already emitted by
// brisk-descriptor-extractor.o), so only declare the one we call.
namespace brisk { void IntegralImage8(const cv::Mat& src, cv::Mat* dest); }

namespace cv {
// The reference declares cv::imread; nothing on the hot path calls it.
Mat imread(const std::string&, int) { return Mat(); }
}  // namespace cv

namespace {
struct RefKeyPoint {  // == cv::KeyPoint field order, 28 bytes
  float x, y, size, angle, response;
  int32_t octave, class_id;
};

cv::Mat WrapCopy(const uint8_t* img, int w, int h) {
  cv::Mat m(h, w, CV_8UC1);
  memcpy(m.data, img, (size_t)w * h);
  return m;
}

void ToVec(const RefKeyPoint* in, int n, std::vector<cv::KeyPoint>* out) {
  out->resize(n);
  for (int i = 0; i < n; ++i) {
    cv::KeyPoint& k = (*out)[i];
    k.pt.x = in[i].x; k.pt.y = in[i].y; k.size = in[i].size; k.angle = in[i].angle;
    k.response = in[i].response; k.octave = in[i].octave; k.class_id = in[i].class_id;
  }
}

int FromVec(const std::vector<cv::KeyPoint>& in, RefKeyPoint* out, int cap) {
  int n = std::min<int>(in.size(), cap);
  for (int i = 0; i < n; ++i) {
    out[i].x = in[i].pt.x; out[i].y = in[i].pt.y; out[i].size = in[i].size;
    out[i].angle = in[i].angle; out[i].response = in[i].response;
    out[i].octave = in[i].octave; out[i].class_id = in[i].class_id;
  }
  return (int)in.size();
}

typedef brisk::ScaleSpaceFeatureDetector<brisk::HarrisScoreCalculator> HarrisDetector;

// Extractors are expensive to build (52 MB look-up table); cache by config.
struct ExtractorKey {
  int rot, scale, version; float pscale;
  bool operator<(const ExtractorKey& o) const {
    if (rot != o.rot) return rot < o.rot;
    if (scale != o.scale) return scale < o.scale;
    if (version != o.version) return version < o.version;
    return pscale < o.pscale;
  }
};
std::mutex g_mu;
std::map<ExtractorKey, std::shared_ptr<brisk::BriskDescriptorExtractor> > g_extractors;

std::shared_ptr<brisk::BriskDescriptorExtractor> GetExtractor(int rot, int scale, int version, float pscale) {
  std::lock_guard<std::mutex> lock(g_mu);
  ExtractorKey key{rot, scale, version, pscale};
  auto it = g_extractors.find(key);
  if (it != g_extractors.end()) return it->second;
  std::shared_ptr<brisk::BriskDescriptorExtractor> e(
      new brisk::BriskDescriptorExtractor(rot != 0, scale != 0, version, pscale));
  g_extractors[key] = e;
  return e;
}
}  // namespace

extern "C" {

int ref_halfsample8(const uint8_t* src, int w, int h, uint8_t* dst) {
  cv::Mat s = WrapCopy(src, w, h);
  cv::Mat d(h / 2, w / 2, CV_8UC1);
  memset(d.data, 0, (size_t)(h / 2) * (w / 2));
  brisk::Halfsample8(s, d);
  memcpy(dst, d.data, (size_t)(h / 2) * (w / 2));
  return 0;
}

// 16-bit samplers (image-down-sampling.cc:56-139,394-548); dst is pre-filled with 0xffff so that pixels the reference does
// not write (images narrower than one SSE block) show up.
int ref_halfsample16(const uint16_t* src, int w, int h, uint16_t* dst) {
  cv::Mat s(h, w, CV_16UC1);
  memcpy(s.data, src, (size_t)w * h * 2);
  cv::Mat d(h / 2, w / 2, CV_16UC1);
  memset(d.data, 0xff, (size_t)(h / 2) * (w / 2) * 2);
  brisk::Halfsample16(s, d);
  memcpy(dst, d.data, (size_t)(h / 2) * (w / 2) * 2);
  return 0;
}

int ref_twothirdsample16(const uint16_t* src, int w, int h, uint16_t* dst) {
  cv::Mat s(h, w, CV_16UC1);
  memcpy(s.data, src, (size_t)w * h * 2);
  const int dw = 2 * (w / 3), dh = 2 * (h / 3);
  cv::Mat d(dh, dw, CV_16UC1);
  memset(d.data, 0xff, (size_t)dh * dw * 2);
  brisk::Twothirdsample16(s, d);
  memcpy(dst, d.data, (size_t)dh * dw * 2);
  return 0;
}

int ref_twothirdsample8(const uint8_t* src, int w, int h, uint8_t* dst) {
  // The SSE loop may read a few bytes past the last row, so
  // give the source slack.
  cv::Mat s(h + 2, w, CV_8UC1);
  memset(s.data, 0, (size_t)(h + 2) * w);
  memcpy(s.data, src, (size_t)w * h);
  s.rows = h;
  const int dw = 2 * (w / 3), dh = 2 * (h / 3);
  cv::Mat d(dh, dw, CV_8UC1);
  memset(d.data, 0, (size_t)dh * dw);
  brisk::Twothirdsample8(s, d);
  memcpy(dst, d.data, (size_t)dh * dw);
  return 0;
}

// One base BriskLayer: threshold map + raw AGAST corners (x, y, score written
// by GetAgastPoints) in detection order.  Returns the corner count.
int ref_layer_dump(const uint8_t* img, int w, int h, int thresh, int lower, uint8_t* thrmap,
                   int32_t* corners_xys, int cap) {
  cv::Mat m = WrapCopy(img, w, h);
  brisk::BriskLayer layer(m, 230, (unsigned char)lower);
  if (thrmap) memcpy(thrmap, layer.thrmap_.data, (size_t)w * h);
  std::vector<cv::KeyPoint> pts;
  layer.GetAgastPoints((uint8_t)thresh, &pts);
  const int n = (int)pts.size();
  for (int i = 0; i < n && i < cap; ++i) {
    const int x = (int)pts[i].pt.x, y = (int)pts[i].pt.y;
    corners_xys[3 * i + 0] = x;
    corners_xys[3 * i + 1] = y;
    corners_xys[3 * i + 2] = layer.scores_.data[x + y * w];
  }
  return n;
}

// Lazy FAST score of one pixel through the reference accessor (fresh layer).
int ref_agast_score(const uint8_t* img, int w, int h, int x, int y, int threshold) {
  cv::Mat m = WrapCopy(img, w, h);
  brisk::BriskLayer layer(m, 230, 10);
  return layer.GetAgastScore(x, y, (uint8_t)threshold);
}

// Dense FAST 9-16 and 5-8 scores at threshold 1 (uncached accessors), for
// pinning the closed-form score.  out916/out58 are w*h bytes.
int ref_dense_scores(const uint8_t* img, int w, int h, uint8_t* out916, uint8_t* out58) {
  cv::Mat m = WrapCopy(img, w, h);
  brisk::BriskLayer layer(m, 230, 10);
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      if (out916) out916[x + y * w] = layer.GetAgastScore(x, y, (uint8_t)1);
      if (out58) out58[x + y * w] = layer.GetAgastScore_5_8(x, y, (uint8_t)1);
    }
  return 0;
}

// Pyramid images, concatenated; dims[2*i]=cols, dims[2*i+1]=rows.
int ref_pyramid(const uint8_t* img, int w, int h, int octaves, uint8_t* out, int32_t* dims, float* scale_offset) {
  cv::Mat m = WrapCopy(img, w, h);
  brisk::BriskScaleSpace ss((uint8_t)octaves, true);
  ss.ConstructPyramid(m, 60);
  size_t off = 0;
  for (size_t i = 0; i < ss.pyramid_.size(); ++i) {
    const cv::Mat& li = ss.pyramid_[i].img();
    dims[2 * i] = li.cols; dims[2 * i + 1] = li.rows;
    if (scale_offset) { scale_offset[2 * i] = ss.pyramid_[i].scale(); scale_offset[2 * i + 1] = ss.pyramid_[i].offset(); }
    if (out) memcpy(out + off, li.data, (size_t)li.cols * li.rows);
    off += (size_t)li.cols * li.rows;
  }
  return (int)ss.pyramid_.size();
}

// Final state of every layer's lazy score cache after GetKeypoints (debug aid
// for the order-dependent cache semantics); layers concatenated.
int ref_agast_cache_dump(const uint8_t* img, int w, int h, int thresh, int octaves, uint8_t* out) {
  cv::Mat m = WrapCopy(img, w, h);
  brisk::BriskScaleSpace ss((uint8_t)octaves, true);
  ss.ConstructPyramid(m, (unsigned char)thresh);
  std::vector<cv::KeyPoint> kps;
  ss.GetKeypoints(&kps);
  size_t off = 0;
  for (size_t i = 0; i < ss.pyramid_.size(); ++i) {
    const cv::Mat& sc = ss.pyramid_[i].scores();
    memcpy(out + off, sc.data, (size_t)sc.cols * sc.rows);
    off += (size_t)sc.cols * sc.rows;
  }
  return (int)kps.size();
}

// brisk::BriskFeatureDetector(thresh, octaves, suppress).detect(image, mask)
int ref_agast_detect(const uint8_t* img, int w, int h, int thresh, int octaves, int suppress,
                     const uint8_t* mask, RefKeyPoint* out, int cap) {
  cv::Mat m = WrapCopy(img, w, h);
  cv::Mat mk;
  if (mask) mk = WrapCopy(mask, w, h);
  brisk::BriskFeatureDetector det(thresh, octaves, suppress != 0);
  std::vector<cv::KeyPoint> kps;
  det.detect(m, kps, mk);
  return FromVec(kps, out, cap);
}

// brisk::BriskFeatureDetector(thresh, octaves, suppress).ComputeScale(image, keypoints): the
// "provided key points" mode of BriskScaleSpace::GetKeypoints (brisk-scale-space.cc:104-124).
int ref_compute_scale(const uint8_t* img, int w, int h, int thresh, int octaves, int suppress,
                      const RefKeyPoint* in, int n_in, RefKeyPoint* out, int cap) {
  cv::Mat m = WrapCopy(img, w, h);
  brisk::BriskFeatureDetector det(thresh, octaves, suppress != 0);
  std::vector<cv::KeyPoint> kps;
  ToVec(in, n_in, &kps);
  det.ComputeScale(m, kps);
  return FromVec(kps, out, cap);
}

// ScaleSpaceFeatureDetector<HarrisScoreCalculator>(octaves, radius, absThr, maxKpt)
int ref_harris_detect(const uint8_t* img, int w, int h, int octaves, double radius, double abs_thr,
                      int64_t max_kpt, RefKeyPoint* out, int cap) {
  cv::Mat m = WrapCopy(img, w, h);
  size_t mk = max_kpt < 0 ? std::numeric_limits<size_t>::max() : (size_t)max_kpt;
  HarrisDetector det((size_t)octaves, radius, abs_thr, mk);
  std::vector<cv::KeyPoint> kps;
  det.detect(m, kps);
  return FromVec(kps, out, cap);
}

// ScaleSpaceFeatureDetector<HarrisScoreCalculator>::detect with a non-empty key-point vector ("use passed key points")
int ref_harris_detect_passed(const uint8_t* img, int w, int h, int octaves, double radius, double abs_thr, int64_t max_kpt,
                             const RefKeyPoint* in, int n_in, RefKeyPoint* out, int cap) {
  cv::Mat m = WrapCopy(img, w, h);
  HarrisDetector det((size_t)octaves, radius, abs_thr, max_kpt < 0 ? std::numeric_limits<size_t>::max() : (size_t)max_kpt);
  std::vector<cv::KeyPoint> kps;
  ToVec(in, n_in, &kps);
  det.detect(m, kps);
  return FromVec(kps, out, cap);
}

// Legacy single-scale brisk::HarrisFeatureDetector(radius) (harris-feature-detector.cc:56-409): detect() on one image.
// scores (nullable) receives the int32 response map of its CornerHarris stage.
int ref_harris_legacy(const uint8_t* img, int w, int h, double radius, int32_t* scores, RefKeyPoint* out, int cap) {
  cv::Mat m = WrapCopy(img, w, h);
  brisk::HarrisFeatureDetector det(radius);
  if (scores) {
    cv::Mat a1, b1, c1, a, b, c, sc;
    brisk::HarrisFeatureDetector::GetCovarEntries(m, a1, b1, c1);
    brisk::FilterGauss3by316S(a1, a);
    brisk::FilterGauss3by316S(b1, b);
    brisk::FilterGauss3by316S(c1, c);
    brisk::HarrisFeatureDetector::CornerHarris(a, b, c, sc);
    memcpy(scores, sc.data, (size_t)w * h * 4);
  }
  std::vector<cv::KeyPoint> kps;
  det.detectImpl(m, kps, cv::Mat());
  return FromVec(kps, out, cap);
}

int ref_harris_scores(const uint8_t* img, int w, int h, int32_t* out) {
  cv::Mat m = WrapCopy(img, w, h);
  cv::Mat scores;
  brisk::HarrisScoresSSE(m, scores);
  memcpy(out, scores.data, (size_t)w * h * 4);
  return 0;
}

// Harris 2-D maxima (raster order) of one image: (score, x, y) triples.
int ref_harris_maxima(const uint8_t* img, int w, int h, int abs_thr, int32_t* out_sxy, int cap) {
  cv::Mat m = WrapCopy(img, w, h);
  brisk::HarrisScoreCalculator calc;
  calc.SetImage(m, true);
  std::vector<brisk::HarrisScoreCalculator::PointWithScore> pts;
  calc.Get2dMaxima(pts, abs_thr);
  const int n = (int)pts.size();
  for (int i = 0; i < n && i < cap; ++i) {
    out_sxy[3 * i] = pts[i].score; out_sxy[3 * i + 1] = pts[i].x; out_sxy[3 * i + 2] = pts[i].y;
  }
  return n;
}

int ref_integral8(const uint8_t* img, int w, int h, int32_t* out) {
  cv::Mat m = WrapCopy(img, w, h);
  cv::Mat integral;
  brisk::IntegralImage8(m, &integral);
  if (integral.rows != h + 1 || integral.cols != w + 1) return -1;
  memcpy(out, integral.data, (size_t)(w + 1) * (h + 1) * 4);
  return 0;
}

// BriskDescriptorExtractor(rot, scale, version, patternScale).compute(); kps is
// in/out (border-culled, angle written).  Returns the surviving count.
int ref_describe(const uint8_t* img, int w, int h, RefKeyPoint* kps, int n, int rot, int scale, int version,
                 float pattern_scale, uint8_t* desc, int32_t* desc_bytes) {
  cv::Mat m = WrapCopy(img, w, h);
  std::shared_ptr<brisk::BriskDescriptorExtractor> ex = GetExtractor(rot, scale, version, pattern_scale);
  std::vector<cv::KeyPoint> v;
  ToVec(kps, n, &v);
  cv::Mat d;
  ex->compute(m, v, d);
  const int nb = ex->descriptorSize();
  if (desc_bytes) *desc_bytes = nb;
  FromVec(v, kps, n);
  if (desc && !v.empty()) memcpy(desc, d.data, (size_t)v.size() * nb);
  return (int)v.size();
}

// Pattern dump for pinning the product's host-side table builder.
// counts: [points, nshort, nlong, strings]; any pointer may be NULL.
int ref_pattern_dump(int version, float pattern_scale, int32_t* counts, float* points_xys /*64*1024*P*3*/,
                     float* scale_list /*64*/, uint32_t* size_list /*64*/, uint32_t* short_pairs /*2*ns*/,
                     int32_t* long_pairs /*4*nl*/) {
  std::shared_ptr<brisk::BriskDescriptorExtractor> ex = GetExtractor(1, 1, version, pattern_scale);
  const unsigned P = ex->points_;
  if (counts) { counts[0] = P; counts[1] = ex->noShortPairs_; counts[2] = ex->noLongPairs_; counts[3] = ex->strings_; }
  if (points_xys) memcpy(points_xys, ex->patternPoints_, sizeof(float) * 3 * P * 64 * 1024);
  if (scale_list) memcpy(scale_list, ex->scaleList_, sizeof(float) * 64);
  if (size_list) memcpy(size_list, ex->sizeList_, sizeof(uint32_t) * 64);
  if (short_pairs)
    for (unsigned p = 0; p < ex->noShortPairs_; ++p) { short_pairs[2 * p] = ex->shortPairs_[p].i; short_pairs[2 * p + 1] = ex->shortPairs_[p].j; }
  if (long_pairs)
    for (unsigned p = 0; p < ex->noLongPairs_; ++p) {
      long_pairs[4 * p] = ex->longPairs_[p].i; long_pairs[4 * p + 1] = ex->longPairs_[p].j;
      long_pairs[4 * p + 2] = ex->longPairs_[p].weighted_dx; long_pairs[4 * p + 3] = ex->longPairs_[p].weighted_dy;
    }
  return 0;
}

// brisk::Hamming()(a, b, nbytes); operands are copied to 16-byte aligned
// scratch because the reference dereferences __m128i* directly.
int ref_hamming(const uint8_t* a, const uint8_t* b, int nbytes) {
  alignas(16) uint8_t aa[256], bb[256];
  if (nbytes > 256 || nbytes % 16) return -1;
  memcpy(aa, a, nbytes); memcpy(bb, b, nbytes);
  brisk::Hamming hm;
  return hm(aa, bb, nbytes);
}

// brisk::BruteForceMatcher itself (brute-force-matcher.cc compiled unmodified against the shim's
// cv::DescriptorMatcher): knnMatch (radius < 0) or radiusMatch over a train collection with optional
// per-image masks [nq][nt_i] (masks == NULL: none; masks[i] == NULL: an empty Mat).  Output, as int32 words:
// number of lists, then per list its length and (queryIdx, trainIdx, imgIdx, distance as float bits) per match.
// Returns the number of words needed (> cap: nothing beyond cap was written).
int64_t ref_matcher(const uint8_t* q, int nq, int nbytes, int n_imgs, const uint8_t* const* trains, const int32_t* nts,
                    const uint8_t* const* masks, int k, float radius, int compact, int32_t* out, int64_t cap) {
  cv::Mat qm(nq, nbytes, CV_8UC1);
  if (nq) memcpy(qm.data, q, (size_t)nq * nbytes);
  std::vector<cv::Mat> coll, mk;
  for (int i = 0; i < n_imgs; ++i) {
    cv::Mat t;
    if (nts[i] > 0) { t.create(nts[i], nbytes, CV_8UC1); memcpy(t.data, trains[i], (size_t)nts[i] * nbytes); }
    coll.push_back(t);
    if (masks) {
      cv::Mat m;
      if (masks[i] && nts[i] > 0) { m.create(nq, nts[i], CV_8UC1); memcpy(m.data, masks[i], (size_t)nq * nts[i]); }
      mk.push_back(m);
    }
  }
  brisk::BruteForceMatcher matcher;
  matcher.add(coll);
  std::vector<std::vector<cv::DMatch> > res;
  if (radius < 0) matcher.knnMatch(qm, res, k, cv::_InputArray(mk), compact != 0);
  else matcher.radiusMatch(qm, res, radius, cv::_InputArray(mk), compact != 0);
  int64_t pos = 0;
  auto put = [&](int32_t v) { if (pos < cap) out[pos] = v; ++pos; };
  put((int32_t)res.size());
  for (const auto& lst : res) {
    put((int32_t)lst.size());
    for (const cv::DMatch& m : lst) {
      int32_t bits; memcpy(&bits, &m.distance, 4);
      put(m.queryIdx); put(m.trainIdx); put(m.imgIdx); put(bits);
    }
  }
  return pos;
}

// Brute-force kNN with the reference distance primitive and the selection rule
// of BruteForceMatcher::commonKnnMatchImpl (k successive arg-mins, first
// minimum wins => lowest train index on ties), multi-threaded over the queries:
// the CPU baseline of the matcher (ref_matcher above is the single-threaded class
// itself).  Rows must be 16-byte
// multiples; q/t need 16-byte alignment (numpy arrays from ctypes are copied).
int ref_knn(const uint8_t* q, int64_t nq, const uint8_t* t, int64_t nt, int nbytes, int k, int32_t* idx,
            int32_t* dist, int nthreads) {
  if (nbytes % 16) return -1;
  uint8_t* qa = (uint8_t*)aligned_alloc(64, ((size_t)nq * nbytes + 63) / 64 * 64 + 64);
  uint8_t* ta = (uint8_t*)aligned_alloc(64, ((size_t)nt * nbytes + 63) / 64 * 64 + 64);
  memcpy(qa, q, (size_t)nq * nbytes); memcpy(ta, t, (size_t)nt * nbytes);
  if (nthreads < 1) nthreads = 1;
  auto work = [&](int tid) {
    brisk::Hamming hm;
    std::vector<int> d(nt);
    for (int64_t i = tid; i < nq; i += nthreads) {
      for (int64_t j = 0; j < nt; ++j) d[j] = hm(qa + i * nbytes, ta + j * nbytes, nbytes);
      for (int kk = 0; kk < k; ++kk) {
        int best = std::numeric_limits<int>::max(); int64_t bj = -1;
        for (int64_t j = 0; j < nt; ++j) if (d[j] < best) { best = d[j]; bj = j; }
        idx[i * k + kk] = (int32_t)bj; dist[i * k + kk] = bj < 0 ? -1 : best;
        if (bj >= 0) d[bj] = std::numeric_limits<int>::max();
      }
    }
  };
  std::vector<std::thread> th;
  for (int i = 0; i < nthreads; ++i) th.emplace_back(work, i);
  for (auto& x : th) x.join();
  free(qa); free(ta);
  return 0;
}

// CPU baseline: BriskFeatureDetector(thresh, octaves) + BriskDescriptorExtractor
// over n frames, frame-parallel over nthreads (one detector + shared const
// extractor per thread).  Returns wall seconds; total_kps receives the number
// of described keypoints.  harris!=0 selects the Harris scale-space detector.
double ref_bench_detect_describe(const uint8_t* imgs, int n, int w, int h, int harris, int thresh, int octaves,
                                 double radius, double abs_thr, int nthreads, int64_t* total_kps) {
  std::shared_ptr<brisk::BriskDescriptorExtractor> ex = GetExtractor(1, 1, 2, 1.0f);
  if (nthreads < 1) nthreads = 1;
  std::vector<int64_t> counts(nthreads, 0);
  auto t0 = std::chrono::steady_clock::now();
  auto work = [&](int tid) {
    brisk::BriskFeatureDetector det(thresh, octaves, true);
    HarrisDetector hdet((size_t)octaves, radius, abs_thr);
    for (int i = tid; i < n; i += nthreads) {
      cv::Mat m = WrapCopy(imgs + (size_t)i * w * h, w, h);
      std::vector<cv::KeyPoint> kps;
      if (harris) hdet.detect(m, kps); else det.detect(m, kps);
      cv::Mat d;
      ex->compute(m, kps, d);
      counts[tid] += (int64_t)kps.size();
    }
  };
  std::vector<std::thread> th;
  for (int i = 0; i < nthreads; ++i) th.emplace_back(work, i);
  for (auto& x : th) x.join();
  auto t1 = std::chrono::steady_clock::now();
  if (total_kps) { *total_kps = 0; for (auto c : counts) *total_kps += c; }
  return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
