// Drop-in C++ surface of ethz-asl/ethzasl_brisk over the B200 C ABI.
//
// Same class names, constructor arguments and call semantics as the reference
// headers (cited per class); every method forwards to libbrisk_b200.so
// (include/brisk_b200.h).  Header-only: link with -lbrisk_b200.  Errors that the
// reference reports with glog CHECK / std::runtime_error are thrown as
// std::runtime_error here.  One GPU context per thread is created on first use
// (the reference classes are likewise not thread-safe per instance).
#ifndef BRISK_BRISK_H_
#define BRISK_BRISK_H_

// With BRISK_B200_USE_OPENCV defined (OpenCV 3 headers available) the classes derive from cv::Feature2D /
// cv::DescriptorMatcher exactly like the reference's (brisk-feature-detector.h:51, brisk-descriptor-extractor.h:54,
// scale-space-feature-detector.h:63, brisk-feature.h:54, brute-force-matcher.h:54), override the same virtuals with the
// same cv::InputArray / cv::OutputArray signatures and the cv:: aliases of brisk/brisk.h:56-59 exist, so a
// cv::Ptr<cv::Feature2D> / cv::Ptr<cv::DescriptorMatcher> holder works unchanged.  Without it the same classes stand
// alone on the minimal agast::Mat / agast::KeyPoint of <agast/wrap-opencv.h> (the reference's own !HAVE_OPENCV branch).
#include <algorithm>
#include <bitset>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include <agast/wrap-opencv.h>
#include "../brisk_b200.h"

#ifdef BRISK_B200_USE_OPENCV
#define BRISK_B200_FEATURE2D : public cv::Feature2D
#else
#define BRISK_B200_FEATURE2D
#endif

namespace brisk {

namespace detail {
inline void check(brisk_ctx* ctx, int rc) {
  if (rc != BRISK_OK) throw std::runtime_error(std::string("brisk_b200: ") + brisk_last_error(ctx));
}
// Per-thread context on device BRISK_B200_DEVICE (default 0).  Shared ownership: every detector / extractor / calculator
// keeps the context it created its handles in alive, so objects with static storage duration, or objects destroyed on
// another thread after their creating thread has exited, never touch a freed context.
typedef std::shared_ptr<brisk_ctx> ContextPtr;
inline ContextPtr context_ptr() {
  static thread_local ContextPtr holder;
  if (!holder) {
    const char* d = std::getenv("BRISK_B200_DEVICE");
    brisk_ctx* ctx = nullptr;
    if (brisk_ctx_create(d ? std::atoi(d) : 0, nullptr, &ctx) != BRISK_OK)
      throw std::runtime_error("brisk_b200: no usable CUDA device (there is no CPU fallback)");
    holder.reset(ctx, brisk_ctx_destroy);
  }
  return holder;
}
inline brisk_ctx* context() { return context_ptr().get(); }
static_assert(sizeof(agast::KeyPoint) == sizeof(brisk_keypoint), "KeyPoint must match cv::KeyPoint's 28-byte layout");

// The C ABI reads a mask with the image's geometry (same size, row stride and frame pitch).  Returns the pointer to pass:
// the mask itself when it is laid out like the image, else a repacked copy in `scratch`; throws on a size mismatch
// (the reference CHECKs the mask size in scale-space-feature-detector.h:113-116 and indexes it by image coordinates in
// brisk-feature-detector.cc:49-66).
inline const unsigned char* mask_like_image(const agast::Mat& image, const agast::Mat& mask, std::vector<unsigned char>* scratch) {
  if (mask.empty()) return nullptr;
  if (mask.rows != image.rows || mask.cols != image.cols) throw std::runtime_error("brisk_b200: mask size differs from the image size");
  if ((size_t)mask.step == (size_t)image.step) return mask.data;
  scratch->assign((size_t)image.step * image.rows, 0);
  for (int r = 0; r < mask.rows; ++r) std::memcpy(scratch->data() + (size_t)r * (size_t)image.step, mask.data + (size_t)r * (size_t)mask.step, mask.cols);
  return scratch->data();
}
}  // namespace detail

// brisk::BriskFeatureDetector -- reference brisk/include/brisk/brisk-feature-detector.h:51-84,
// brisk/src/brisk-feature-detector.cc:69-92.
class BriskFeatureDetector BRISK_B200_FEATURE2D {
 public:
  BriskFeatureDetector(int thresh, int octaves_ = 3, bool suppressScaleNonmaxima = true)
      : threshold(thresh), octaves(octaves_), m_suppress(suppressScaleNonmaxima) {}
  virtual ~BriskFeatureDetector() { if (det_) brisk_detector_destroy(det_); }
  BriskFeatureDetector(const BriskFeatureDetector&) = delete;
  BriskFeatureDetector& operator=(const BriskFeatureDetector&) = delete;

  int threshold;
  int octaves;

#ifdef BRISK_B200_USE_OPENCV
  // cv::Feature2D::detect() forwards here; the reference's override (brisk-feature-detector.h:69-74): detection only
  virtual void detectAndCompute(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint>& keypoints,
                                cv::OutputArray /*descriptors*/, bool /*useProvidedKeypoints*/ = false) {
    detectImpl(image.getMat(), keypoints, mask.getMat());
  }
#else
  // cv::Feature2D::detect
  void detect(const agast::Mat& image, std::vector<agast::KeyPoint>& keypoints, const agast::Mat& mask = agast::Mat()) const {
    detectImpl(image, keypoints, mask);
  }
  // cv::Feature2D::detectAndCompute as the reference overrides it (brisk-feature-detector.h:69-74): detection only
  virtual void detectAndCompute(const agast::Mat& image, const agast::Mat& mask, std::vector<agast::KeyPoint>& keypoints,
                                agast::Mat& /*descriptors*/, bool /*useProvidedKeypoints*/ = false) {
    detectImpl(image, keypoints, mask);
  }
#endif
  // brisk-feature-detector.cc:87-92: re-examines the passed key points in every pyramid layer and replaces
  // them by the key points the scale-space checks accept (one per accepting layer; octave = layer index).
  // A layer that keeps none of the points is detected on instead, as in the reference (brisk-layer.cc:103-105).
  // With an empty vector the reference detects on every layer (threshold map without lower bound): that case
  // is reported as std::runtime_error here -- call detect().
  void ComputeScale(const agast::Mat& image, std::vector<agast::KeyPoint>& keypoints) const {
    brisk_ctx* ctx = ensure();
    const int32_t n_in = (int32_t)keypoints.size();
    const int n_layers = octaves == 0 ? 1 : 2 * octaves;
    std::vector<agast::KeyPoint> out((size_t)std::max(1, n_layers * n_in));
    for (;;) {
      int32_t count = 0;
      const int rc = brisk_compute_scale(ctx, det_, image.data, 1, image.cols, image.rows, image.step, image.step * image.rows,
                                         reinterpret_cast<const brisk_keypoint*>(keypoints.data()), &n_in, std::max(1, n_in),
                                         reinterpret_cast<brisk_keypoint*>(out.data()), &count, (int)out.size());
      // layers without points contribute their corners: grow and retry
      if (rc == BRISK_ERR_CAPACITY && count > (int)out.size()) { out.resize(count); continue; }
      detail::check(ctx, rc);
      out.resize(count);
      keypoints.swap(out);
      return;
    }
  }
  // raw-corner capacity per frame (B200 specific; default scales with the image area)
  void setCornerCapacity(int n) { corner_cap_ = n; if (det_) brisk_detector_set_corner_capacity(det_, n); }

 protected:
  virtual void detectImpl(const agast::Mat& image, std::vector<agast::KeyPoint>& keypoints, const agast::Mat& mask) const {
    keypoints.clear();
    if (image.empty()) return;
    brisk_ctx* ctx = ensure();
    std::vector<unsigned char> mask_scratch;
    const unsigned char* mask_ptr = detail::mask_like_image(image, mask, &mask_scratch);
    int cap = (int)std::max<long long>(4096, (long long)image.rows * image.cols / 64);
    for (;;) {
      keypoints.resize(cap);
      int32_t count = 0;
      const int rc = brisk_detect(ctx, det_, image.data, 1, image.cols, image.rows, image.step, image.step * image.rows,
                                  mask_ptr, reinterpret_cast<brisk_keypoint*>(keypoints.data()), &count, cap);
      if (rc == BRISK_ERR_CAPACITY && count > cap) { cap = count; continue; }  // grow and retry
      detail::check(ctx, rc);
      keypoints.resize(count);
      return;
    }
  }

 private:
  // (re)creates the detector handle in the calling thread's context when the public parameters changed; the handle's
  // context is kept alive by ctx_
  brisk_ctx* ensure() const {
    detail::ContextPtr cur = detail::context_ptr();
    if (det_ && ctx_ == cur && cfg_thr_ == threshold && cfg_oct_ == octaves) return cur.get();
    if (det_) brisk_detector_destroy(det_);
    det_ = nullptr;
    detail::check(cur.get(), brisk_agast_detector_create(cur.get(), threshold, octaves, m_suppress ? 1 : 0, &det_));
    if (corner_cap_ > 0) brisk_detector_set_corner_capacity(det_, corner_cap_);
    ctx_ = cur; cfg_thr_ = threshold; cfg_oct_ = octaves;
    return cur.get();
  }
  bool m_suppress;
  int corner_cap_ = 0;
  mutable detail::ContextPtr ctx_;   // declared before det_'s users: released after the destructor body destroyed det_
  mutable brisk_detector* det_ = nullptr;
  mutable int cfg_thr_ = 0, cfg_oct_ = 0;
};

// brisk::HarrisScoreCalculator : ScoreCalculator<int> -- reference brisk/include/brisk/harris-score-calculator.h:52-90,
// internal/score-calculator.h:60-127.  SetImage computes the integer Harris score map (HarrisScoresSSE) on the GPU and keeps
// it on the host; Score(int, int) / Score(double, double) are the reference's inline accessors on that map; Get2dMaxima
// lists the 8-neighbour maxima >= absoluteThreshold in raster order.  Also the template argument of
// ScaleSpaceFeatureDetector, which is the only use the reference itself makes of it.
class HarrisScoreCalculator {
 public:
  typedef int Score_t;
  struct PointWithScore {  // score-calculator.h:69-91 (USE_SIMPLE_POINT_WITH_SCORE)
    PointWithScore() : score(0), x(0), y(0) {}
    PointWithScore(Score_t score_, uint16_t x_, uint16_t y_) : score(score_), x(x_), y(y_) {}
    Score_t score;
    uint16_t x, y;
    bool operator<(const PointWithScore& other) const { return score > other.score; }  // (sic) sorts descending
  };
  virtual ~HarrisScoreCalculator() {}

  void SetImage(const agast::Mat& img, bool initScores = true) {
    _img = img;
    if (initScores) InitializeScores();
  }
  inline double Score(double u, double v) {
    const int u_int = static_cast<int>(u), v_int = static_cast<int>(v);
    if (u_int + 1 >= cols_ || v_int + 1 >= rows_ || u_int < 0 || v_int < 0) return 0.0;
    const double ru = u - static_cast<double>(u_int), rv = v - static_cast<double>(v_int);
    const double oneMinus_ru = 1.0 - ru, oneMinus_rv = 1.0 - rv;
    return oneMinus_rv * (oneMinus_ru * at(v_int, u_int) + ru * at(v_int, u_int + 1)) +
           rv * (oneMinus_ru * at(v_int + 1, u_int) + ru * at(v_int + 1, u_int + 1));
  }
  inline Score_t Score(int u, int v) { return at(v, u); }
  virtual void Get2dMaxima(std::vector<PointWithScore>& points, Score_t absoluteThreshold = 0) {
    if (_img.empty()) return;
    detail::ContextPtr ctx = detail::context_ptr();
    int cap = std::max(4000, _img.rows * _img.cols / 32);
    for (;;) {
      std::vector<int32_t> sxy((size_t)cap * 3);
      int32_t n = 0;
      const int rc = brisk_harris_scores(ctx.get(), _img.data, _img.cols, _img.rows, _img.step, absoluteThreshold, nullptr, sxy.data(), cap, &n);
      if (rc == BRISK_ERR_CAPACITY && n > cap) { cap = n; continue; }
      detail::check(ctx.get(), rc);
      points.reserve(points.size() + (size_t)n);
      for (int i = 0; i < n; ++i) points.push_back(PointWithScore(sxy[3 * i], (uint16_t)sxy[3 * i + 1], (uint16_t)sxy[3 * i + 2]));
      return;
    }
  }

 protected:
  virtual void InitializeScores() {
    detail::ContextPtr ctx = detail::context_ptr();
    rows_ = _img.rows; cols_ = _img.cols;
    _scores.assign((size_t)rows_ * cols_, 0);
    if (_img.empty()) return;
    detail::check(ctx.get(), brisk_harris_scores(ctx.get(), _img.data, _img.cols, _img.rows, _img.step, 0, _scores.data(), nullptr, 0, nullptr));
  }
  int at(int row, int col) const { return _scores[(size_t)row * cols_ + col]; }
  agast::Mat _img;             // the image we operate on
  std::vector<int> _scores;    // calculated scores, rows_ x cols_
  int rows_ = 0, cols_ = 0;
};

// brisk::ScaleSpaceFeatureDetector<SCORE_CALCULATOR_T> -- reference
// brisk/include/brisk/scale-space-feature-detector.h:62-135.  Empty images return silently; the
// mask is ignored (as in the reference's detectImpl); a non-empty input vector switches to the "use passed
// key points" mode (:103-108): no detection, the passed points with response > 1e6 are re-filtered by the
// uniformity enforcement / bucketing and returned unrefined (defined for octaves == 0 only).
template <class SCORE_CALCULATOR_T>
class ScaleSpaceFeatureDetector BRISK_B200_FEATURE2D {
 public:
  typedef SCORE_CALCULATOR_T ScoreCalculator_t;
  ScaleSpaceFeatureDetector(size_t octaves, double uniformityRadius, double absoluteThreshold = 0,
                            size_t maxNumKpt = std::numeric_limits<size_t>::max())
      : _octaves(octaves), _uniformityRadius(uniformityRadius), _absoluteThreshold(absoluteThreshold), _maxNumKpt(maxNumKpt) {}
  virtual ~ScaleSpaceFeatureDetector() { if (det_) brisk_detector_destroy(det_); }
  ScaleSpaceFeatureDetector(const ScaleSpaceFeatureDetector&) = delete;
  ScaleSpaceFeatureDetector& operator=(const ScaleSpaceFeatureDetector&) = delete;

  void detect(const agast::Mat& image, std::vector<agast::KeyPoint>& keypoints, const agast::Mat& mask = agast::Mat()) const {
    detectImpl(image, keypoints, mask);
  }

#ifdef BRISK_B200_USE_OPENCV
  // scale-space-feature-detector.h:92-97: detection only
  virtual void detectAndCompute(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint>& keypoints,
                                cv::OutputArray /*descriptors*/, bool /*useProvidedKeypoints*/ = false) {
    detectImpl(image.getMat(), keypoints, mask.getMat());
  }
#else
  virtual void detectAndCompute(const agast::Mat& image, const agast::Mat& mask, std::vector<agast::KeyPoint>& keypoints,
                                agast::Mat& /*descriptors*/, bool /*useProvidedKeypoints*/ = false) {
    detectImpl(image, keypoints, mask);
  }
#endif

 protected:
  // scale-space-feature-detector.h:100-128
  virtual void detectImpl(const agast::Mat& image, std::vector<agast::KeyPoint>& keypoints, const agast::Mat& mask = agast::Mat()) const {
    if (image.empty()) return;
    if (!mask.empty() && (mask.rows != image.rows || mask.cols != image.cols))  // the reference CHECKs this (:113-116); the mask itself is unused
      throw std::runtime_error("brisk_b200: mask size differs from the image size");
    detail::ContextPtr cur = detail::context_ptr();
    brisk_ctx* ctx = cur.get();
    if (!det_ || ctx_ != cur) {
      if (det_) brisk_detector_destroy(det_);
      det_ = nullptr;
      const int64_t mk = _maxNumKpt > (size_t)std::numeric_limits<int64_t>::max() ? -1 : (int64_t)_maxNumKpt;
      detail::check(ctx, brisk_harris_detector_create(ctx, (int)_octaves, _uniformityRadius, _absoluteThreshold, mk, &det_));
      ctx_ = cur;
    }
    if (!keypoints.empty()) {  // use the passed key points
      const int32_t n_in = (int32_t)keypoints.size();
      std::vector<agast::KeyPoint> out(keypoints.size());
      int32_t count = 0;
      detail::check(ctx, brisk_harris_detect_passed(ctx, det_, 1, image.cols, image.rows, reinterpret_cast<const brisk_keypoint*>(keypoints.data()),
                                                    &n_in, n_in, reinterpret_cast<brisk_keypoint*>(out.data()), &count, n_in));
      out.resize(count);
      keypoints.swap(out);
      return;
    }
    int cap = (int)std::max<long long>(4096, (long long)image.rows * image.cols / 64);
    for (;;) {
      keypoints.resize(cap);
      int32_t count = 0;
      const int rc = brisk_detect(ctx, det_, image.data, 1, image.cols, image.rows, image.step, image.step * image.rows, nullptr,
                                  reinterpret_cast<brisk_keypoint*>(keypoints.data()), &count, cap);
      if (rc == BRISK_ERR_CAPACITY && count > cap) { cap = count; continue; }
      detail::check(ctx, rc);
      keypoints.resize(count);
      return;
    }
  }

  size_t _octaves;
  double _uniformityRadius;
  double _absoluteThreshold;
  size_t _maxNumKpt;

 private:
  mutable detail::ContextPtr ctx_;
  mutable brisk_detector* det_ = nullptr;
};
typedef ScaleSpaceFeatureDetector<HarrisScoreCalculator> HarrisScaleSpaceFeatureDetector;

// brisk::HarrisFeatureDetector -- the legacy single-scale detector, reference brisk/include/brisk/harris-feature-detector.h:
// 51-82, brisk/src/harris-feature-detector.cc:56-409.  The mask is ignored, as in the reference's detectImpl.  Images for which
// the reference's transposed occupancy indexing leaves its map (landscape shapes) throw std::runtime_error.
class HarrisFeatureDetector BRISK_B200_FEATURE2D {
 public:
  explicit HarrisFeatureDetector(double radius) { SetRadius(radius); }
  virtual ~HarrisFeatureDetector() { if (det_) brisk_detector_destroy(det_); }
  HarrisFeatureDetector(const HarrisFeatureDetector&) = delete;
  HarrisFeatureDetector& operator=(const HarrisFeatureDetector&) = delete;
  void SetRadius(double radius) { _radius = radius; if (det_) { brisk_detector_destroy(det_); det_ = nullptr; } }
#ifdef BRISK_B200_USE_OPENCV
  virtual void detectAndCompute(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint>& keypoints,
                                cv::OutputArray /*descriptors*/, bool /*useProvidedKeypoints*/ = false) {
    detectImpl(image.getMat(), keypoints, mask.getMat());
  }
#else
  void detect(const agast::Mat& image, std::vector<agast::KeyPoint>& keypoints, const agast::Mat& mask = agast::Mat()) const {
    detectImpl(image, keypoints, mask);
  }
#endif

 protected:
  virtual void detectImpl(const agast::Mat& image, std::vector<agast::KeyPoint>& keypoints, const agast::Mat& /*mask*/ = agast::Mat()) const {
    keypoints.resize(0);
    if (image.empty()) return;
    detail::ContextPtr cur = detail::context_ptr();
    brisk_ctx* ctx = cur.get();
    if (!det_ || ctx_ != cur) {
      if (det_) brisk_detector_destroy(det_);
      det_ = nullptr;
      detail::check(ctx, brisk_harris_legacy_detector_create(ctx, _radius, &det_));
      ctx_ = cur;
    }
    int cap = (int)std::max<long long>(4096, (long long)image.rows * image.cols / 64);
    for (;;) {
      keypoints.resize(cap);
      int32_t count = 0;
      const int rc = brisk_detect(ctx, det_, image.data, 1, image.cols, image.rows, image.step, image.step * image.rows, nullptr,
                                  reinterpret_cast<brisk_keypoint*>(keypoints.data()), &count, cap);
      if (rc == BRISK_ERR_CAPACITY && count > cap) { cap = count; continue; }
      detail::check(ctx, rc);
      keypoints.resize(count);
      return;
    }
  }
  double _radius;

 private:
  mutable detail::ContextPtr ctx_;
  mutable brisk_detector* det_ = nullptr;
};

// brisk::BriskDescriptorExtractor -- reference brisk/include/brisk/brisk-descriptor-extractor.h:54-202.
class BriskDescriptorExtractor BRISK_B200_FEATURE2D {
 public:
  static const unsigned int kDescriptorLength = 384;
  enum Version { briskV1 = 1, briskV2 = 2 };

  BriskDescriptorExtractor() : BriskDescriptorExtractor(true, true) {}
  BriskDescriptorExtractor(bool rotationInvariant, bool scaleInvariant) : BriskDescriptorExtractor(rotationInvariant, scaleInvariant, briskV2, 1.0f) {}
  BriskDescriptorExtractor(bool rotationInvariant, bool scaleInvariant, int version) : BriskDescriptorExtractor(rotationInvariant, scaleInvariant, version, 1.0f) {}
  BriskDescriptorExtractor(bool rotationInvariant, bool scaleInvariant, int version, float patternScale)
      : rotationInvariance(rotationInvariant), scaleInvariance(scaleInvariant), version_(version), pattern_scale_(patternScale) {
    if (version != briskV1 && version != briskV2) throw std::runtime_error("only Version::briskV1 or Version::briskV2 supported!");
  }
  explicit BriskDescriptorExtractor(const std::string& fname, bool rotationInvariant = true, bool scaleInvariant = true, float patternScale = 1.0f)
      : rotationInvariance(rotationInvariant), scaleInvariance(scaleInvariant), version_(briskV2), pattern_scale_(patternScale), fname_(fname) {}
  virtual ~BriskDescriptorExtractor() { if (ext_) brisk_extractor_destroy(ext_); }
  BriskDescriptorExtractor(const BriskDescriptorExtractor&) = delete;
  BriskDescriptorExtractor& operator=(const BriskDescriptorExtractor&) = delete;

  bool rotationInvariance;
  bool scaleInvariance;

  int descriptorSize() const { ensure(); return brisk_extractor_descriptor_size(ext_); }
  int descriptorType() const { return CV_8U; }

  // compute(): removes key points too close to the border, writes their angle, fills an N x descriptorSize() matrix
  virtual void compute(const agast::Mat& image, std::vector<agast::KeyPoint>& keypoints, agast::Mat& descriptors) const {
    computeImpl(image, keypoints, descriptors);
  }
  virtual void compute(const agast::Mat& image, std::vector<agast::KeyPoint>& keypoints,
                       std::vector<std::bitset<kDescriptorLength> >& descriptors) const {
    computeImpl(image, keypoints, descriptors);
  }
#ifdef BRISK_B200_USE_OPENCV
  // brisk-descriptor-extractor.h:120-125: extraction only
  virtual void detectAndCompute(cv::InputArray image, cv::InputArray /*mask*/, std::vector<cv::KeyPoint>& keypoints,
                                cv::OutputArray descriptors, bool /*useProvidedKeypoints*/ = false) {
    computeImpl(image.getMat(), keypoints, descriptors.getMatRef());
  }
#else
  virtual void detectAndCompute(const agast::Mat& image, const agast::Mat& /*mask*/, std::vector<agast::KeyPoint>& keypoints,
                                agast::Mat& descriptors, bool /*useProvidedKeypoints*/ = false) {
    computeImpl(image, keypoints, descriptors);
  }
#endif

 protected:
  // brisk-descriptor-extractor.cc:589-599
  virtual void computeImpl(const agast::Mat& image, std::vector<agast::KeyPoint>& keypoints, agast::Mat& descriptors) const {
    brisk_ctx* ctx = ensure();
    const int nb = brisk_extractor_descriptor_size(ext_);
    int32_t count = (int32_t)keypoints.size();
    const int cap = std::max<int>(count, 1);
    std::vector<unsigned char> buf((size_t)cap * nb);
    keypoints.resize(cap);
    detail::check(ctx, brisk_describe(ctx, ext_, image.data, 1, image.cols, image.rows, image.step, image.step * image.rows,
                                      reinterpret_cast<brisk_keypoint*>(keypoints.data()), &count, cap, buf.data()));
    keypoints.resize(count);
    descriptors = agast::Mat::zeros(count, nb, CV_8UC1);
    if (count) std::memcpy(descriptors.data, buf.data(), (size_t)count * nb);
  }
  virtual void computeImpl(const agast::Mat& image, std::vector<agast::KeyPoint>& keypoints,
                           std::vector<std::bitset<kDescriptorLength> >& descriptors) const {
    agast::Mat d;
    computeImpl(image, keypoints, d);
    descriptors.assign(keypoints.size(), std::bitset<kDescriptorLength>());
    for (size_t k = 0; k < keypoints.size(); ++k)
      for (unsigned p = 0; p < kDescriptorLength && p < 8u * d.cols; ++p)
        if (d.data[k * d.cols + p / 8] >> (p % 8) & 1) descriptors[k].set(p, true);
  }

 private:
  brisk_ctx* ensure() const {
    detail::ContextPtr cur = detail::context_ptr();
    if (ext_ && ctx_ == cur && cfg_rot_ == rotationInvariance && cfg_scale_ == scaleInvariance) return cur.get();
    if (ext_) brisk_extractor_destroy(ext_);
    ext_ = nullptr;
    detail::check(cur.get(), brisk_extractor_create(cur.get(), rotationInvariance, scaleInvariance, version_, pattern_scale_,
                                                    fname_.empty() ? nullptr : fname_.c_str(), &ext_));
    ctx_ = cur; cfg_rot_ = rotationInvariance; cfg_scale_ = scaleInvariance;
    return cur.get();
  }
  int version_;
  float pattern_scale_;
  std::string fname_;
  mutable detail::ContextPtr ctx_;
  mutable brisk_extractor* ext_ = nullptr;
  mutable bool cfg_rot_ = false, cfg_scale_ = false;
};

// brisk::BriskFeature -- reference brisk/include/brisk/brisk-feature.h:54-114: Harris scale-space
// detector + BRISK extractor behind one detectAndCompute().
class BriskFeature BRISK_B200_FEATURE2D {
 public:
  BriskFeature(size_t octaves, double uniformityRadius, double absoluteThreshold = 0,
               size_t maxNumKpt = std::numeric_limits<size_t>::max(), bool rotationInvariant = true, bool scaleInvariant = true,
               int extractorVersion = BriskDescriptorExtractor::briskV2)
      : _briskDetector(octaves, uniformityRadius, absoluteThreshold, maxNumKpt),
        _briskExtractor(rotationInvariant, scaleInvariant, extractorVersion) {}
  int descriptorSize() const { return _briskExtractor.descriptorSize(); }
  int descriptorType() const { return _briskExtractor.descriptorType(); }
#ifdef BRISK_B200_USE_OPENCV
  virtual void detectAndCompute(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint>& keypoints,
                                cv::OutputArray descriptors, bool useProvidedKeypoints = false) {
    run(image.getMat(), mask.getMat(), keypoints, descriptors.getMatRef(), useProvidedKeypoints);
  }
#else
  virtual void detectAndCompute(const agast::Mat& image, const agast::Mat& mask, std::vector<agast::KeyPoint>& keypoints,
                                agast::Mat& descriptors, bool useProvidedKeypoints = false) {
    run(image, mask, keypoints, descriptors, useProvidedKeypoints);
  }
#endif
  virtual ~BriskFeature() {}

 private:
  void run(const agast::Mat& image, const agast::Mat& mask, std::vector<agast::KeyPoint>& keypoints, agast::Mat& descriptors,
           bool useProvidedKeypoints) {
    // brisk-feature.h:80-93: the detector runs either way -- on provided key points it re-filters them
    // ("use passed key points") instead of detecting
    if (!useProvidedKeypoints) keypoints.clear();
    _briskDetector.detect(image, keypoints, mask);
    _briskExtractor.compute(image, keypoints, descriptors);
  }
  ScaleSpaceFeatureDetector<HarrisScoreCalculator> _briskDetector;
  BriskDescriptorExtractor _briskExtractor;
};

// brisk::Hamming -- reference brisk/include/brisk/internal/hamming.h:56-114.
// NOTE for callers that loop over descriptor pairs (OKVIS-style): one operator() call is one host -> device -> host round
// trip (~10 us), four orders of magnitude slower than the SSE primitive it replaces.  Use Hamming::distances() (n pairs per
// call, brisk_hamming_distance) or BruteForceMatcher (all pairs on the device) instead; operator() exists for API parity.
class Hamming {
 public:
  typedef unsigned char ValueType;
  typedef int ResultType;
  // n row pairs in one call: out[i] = popcount(a_i xor b_i) over size / 16 whole 128-bit words (host or device pointers)
  static void distances(const unsigned char* a, const unsigned char* b, int64_t n, int size, int32_t* out) {
    brisk_ctx* ctx = detail::context();
    detail::check(ctx, brisk_hamming_distance(ctx, a, b, n, size, out));
  }
  ResultType operator()(const unsigned char* a, const unsigned char* b, int size) const {
    int32_t d = 0;
    brisk_ctx* ctx = detail::context();
    detail::check(ctx, brisk_hamming_distance(ctx, a, b, 1, size, &d));
    return d;
  }
  // hamming.h:79-91: the distance over `numberOf128BitWords` 16-byte words
  static uint32_t PopcntofXORed(const unsigned char* signature1, const unsigned char* signature2, const int numberOf128BitWords) {
    int32_t d = 0;
    brisk_ctx* ctx = detail::context();
    detail::check(ctx, brisk_hamming_distance(ctx, signature1, signature2, 1, 16 * numberOf128BitWords, &d));
    return (uint32_t)d;
  }
};

#ifdef BRISK_B200_USE_OPENCV
typedef cv::DMatch DMatch;
#else
struct DMatch {  // == cv::DMatch
  int queryIdx = -1, trainIdx = -1, imgIdx = -1;
  float distance = std::numeric_limits<float>::max();
  bool operator<(const DMatch& m) const { return distance < m.distance; }
};
#endif

// brisk::BruteForceMatcher -- reference brisk/include/brisk/brute-force-matcher.h:54-94 and
// brisk/src/brute-force-matcher.cc:59-214, with the cv::DescriptorMatcher conventions it relies on: a train
// collection (add / clear), per-image masks (query rows x train rows, 0 = pair excluded; an empty mask allows
// everything), compactResult, knnMatch / radiusMatch / match.  Distances and candidate selection run on the GPU;
// this class only reshapes the results into the reference's DMatch lists.
#ifdef BRISK_B200_USE_OPENCV
class BruteForceMatcher : public cv::DescriptorMatcher {
 public:
  explicit BruteForceMatcher(const Hamming& = Hamming()) {}
  virtual ~BruteForceMatcher() {}
  virtual bool isMaskSupported() const { return true; }
  virtual cv::Ptr<cv::DescriptorMatcher> clone(bool emptyTrainData = false) const {
    BruteForceMatcher* m = new BruteForceMatcher();
    if (!emptyTrainData)
      for (const auto& d : trainDescCollection) m->trainDescCollection.push_back(d.clone());
    return cv::Ptr<cv::DescriptorMatcher>(m);
  }

 protected:
  // brute-force-matcher.h:66-75; cv::DescriptorMatcher::knnMatch / radiusMatch / match forward here
  virtual void knnMatchImpl(cv::InputArray queryDescriptors, std::vector<std::vector<cv::DMatch> >& matches, int k,
                            cv::InputArrayOfArrays masks = cv::noArray(), bool compactResult = false) {
    std::vector<cv::Mat> mv;
    masks.getMatVector(mv);
    knnImpl(queryDescriptors.getMat(), trainDescCollection, matches, k, mv, compactResult);
  }
  virtual void radiusMatchImpl(cv::InputArray queryDescriptors, std::vector<std::vector<cv::DMatch> >& matches, float maxDistance,
                               cv::InputArrayOfArrays masks = cv::noArray(), bool compactResult = false) {
    std::vector<cv::Mat> mv;
    masks.getMatVector(mv);
    radiusImpl(queryDescriptors.getMat(), trainDescCollection, matches, maxDistance, mv, compactResult);
  }
#else
class BruteForceMatcher {
 public:
  explicit BruteForceMatcher(const Hamming& = Hamming()) {}
  bool isMaskSupported() const { return true; }
  void add(const std::vector<agast::Mat>& descriptors) { for (const auto& d : descriptors) train_.push_back(d); }
  void clear() { train_.clear(); }
  bool empty() const { return train_.empty(); }
  const std::vector<agast::Mat>& getTrainDescriptors() const { return train_; }
  BruteForceMatcher clone(bool emptyTrainData = false) const { BruteForceMatcher m; if (!emptyTrainData) m.train_ = train_; return m; }

  void knnMatch(const agast::Mat& query, const agast::Mat& train, std::vector<std::vector<DMatch> >& matches, int k,
                const agast::Mat& mask = agast::Mat(), bool compactResult = false) const {
    knnImpl(query, std::vector<agast::Mat>(1, train), matches, k, mask.empty() ? std::vector<agast::Mat>() : std::vector<agast::Mat>(1, mask), compactResult);
  }
  void knnMatch(const agast::Mat& query, std::vector<std::vector<DMatch> >& matches, int k,
                const std::vector<agast::Mat>& masks = std::vector<agast::Mat>(), bool compactResult = false) const {
    knnImpl(query, train_, matches, k, masks, compactResult);
  }
  void radiusMatch(const agast::Mat& query, const agast::Mat& train, std::vector<std::vector<DMatch> >& matches, float maxDistance,
                   const agast::Mat& mask = agast::Mat(), bool compactResult = false) const {
    radiusImpl(query, std::vector<agast::Mat>(1, train), matches, maxDistance, mask.empty() ? std::vector<agast::Mat>() : std::vector<agast::Mat>(1, mask), compactResult);
  }
  void radiusMatch(const agast::Mat& query, std::vector<std::vector<DMatch> >& matches, float maxDistance,
                   const std::vector<agast::Mat>& masks = std::vector<agast::Mat>(), bool compactResult = false) const {
    radiusImpl(query, train_, matches, maxDistance, masks, compactResult);
  }
  // cv::DescriptorMatcher::match: knnMatch(k = 1, compactResult = true), flattened
  void match(const agast::Mat& query, const agast::Mat& train, std::vector<DMatch>& matches, const agast::Mat& mask = agast::Mat()) const {
    std::vector<std::vector<DMatch> > knn;
    knnMatch(query, train, knn, 1, mask, true);
    matches.clear();
    for (const auto& v : knn) matches.insert(matches.end(), v.begin(), v.end());
  }
#endif  // BRISK_B200_USE_OPENCV

 private:
  struct Collection {
    std::vector<unsigned char> all, mask;
    std::vector<int> start;
    std::vector<char> masked_out;
    int last_nonempty = -1;
  };
  // concatenated train rows (global row -> (imgIdx, trainIdx) through `start`), masks side by side, and
  // cv::DescriptorMatcher::isMaskedOut per query
  static void gather(const agast::Mat& query, const std::vector<agast::Mat>& train, const std::vector<agast::Mat>& masks, Collection* c) {
    c->start.assign(1, 0);
    for (size_t i = 0; i < train.size(); ++i) {
      const agast::Mat& t = train[i];
      if (!t.empty() && t.cols != query.cols) throw std::runtime_error("descriptor size mismatch");
      for (int r = 0; r < t.rows; ++r) c->all.insert(c->all.end(), t.data + (size_t)r * t.step, t.data + (size_t)r * t.step + t.cols);
      c->start.push_back(c->start.back() + t.rows);
      if (t.rows > 0) c->last_nonempty = (int)i;
    }
    c->masked_out.assign(query.rows, 0);
    if (masks.empty()) return;
    if (masks.size() != train.size()) throw std::runtime_error("one mask per train image expected");
    const int nt = c->start.back();
    c->mask.assign((size_t)query.rows * nt, 1);
    bool every_mask_given = true;
    for (size_t i = 0; i < masks.size(); ++i) {
      const agast::Mat& m = masks[i];
      if (m.empty()) { every_mask_given = false; continue; }
      if (m.rows != query.rows || m.cols != train[i].rows) throw std::runtime_error("mask size mismatch");
      for (int q = 0; q < query.rows; ++q)
        for (int t = 0; t < m.cols; ++t) c->mask[(size_t)q * nt + c->start[i] + t] = m.data[(size_t)q * m.step + t] != 0;
    }
    if (every_mask_given)
      for (int q = 0; q < query.rows; ++q) {
        bool any = false;
        for (int t = 0; t < nt && !any; ++t) any = c->mask[(size_t)q * nt + t] != 0;
        c->masked_out[q] = !any;
      }
  }
  static std::vector<unsigned char> tight(const agast::Mat& m) {
    std::vector<unsigned char> v((size_t)m.rows * m.cols);
    for (int r = 0; r < m.rows; ++r) std::memcpy(v.data() + (size_t)r * m.cols, m.data + (size_t)r * m.step, m.cols);
    return v;
  }
  static DMatch make(int q, int g, const std::vector<int>& start, float d) {
    int img = 0;
    while (g >= start[img + 1]) ++img;
    DMatch m;
    m.queryIdx = q; m.trainIdx = g - start[img]; m.imgIdx = img; m.distance = d;
    return m;
  }
  void knnImpl(const agast::Mat& query, const std::vector<agast::Mat>& train, std::vector<std::vector<DMatch> >& matches, int k,
               const std::vector<agast::Mat>& masks, bool compactResult) const {
    matches.clear();
    if (query.rows == 0) return;
    Collection c;
    gather(query, train, masks, &c);
    const int nt = c.start.back();
    std::vector<int32_t> idx((size_t)query.rows * k, -1), dist((size_t)query.rows * k, -1);
    if (nt > 0) {
      const std::vector<unsigned char> q = tight(query);
      brisk_ctx* ctx = detail::context();
      detail::check(ctx, brisk_hamming_knn_masked(ctx, q.data(), query.rows, c.all.data(), nt, query.cols, k,
                                                  c.mask.empty() ? nullptr : c.mask.data(), idx.data(), dist.data()));
    }
    for (int q = 0; q < query.rows; ++q) {
      if (c.masked_out[q]) { if (!compactResult) matches.push_back(std::vector<DMatch>()); continue; }
      matches.push_back(std::vector<DMatch>());
      if (nt == 0) continue;
      for (int j = 0; j < k; ++j) {
        const int g = idx[(size_t)q * k + j];
        if (g >= 0) matches.back().push_back(make(q, g, c.start, (float)dist[(size_t)q * k + j]));
        else {
          // The reference keeps selecting once the real candidates have run out (brute-force-matcher.cc:138-157):
          // every entry then holds INT_MAX, minMaxLoc returns location 0, and INT_MAX as a double is below the
          // float it was just rounded to, so the last non-empty image wins.
          DMatch m;
          m.queryIdx = q; m.trainIdx = 0; m.imgIdx = c.last_nonempty; m.distance = 2147483648.0f;
          matches.back().push_back(m);
        }
      }
      // the reference's final std::sort (brute-force-matcher.cc:160): the list is in ascending order already, which
      // insertion sort (up to 16 entries) keeps; longer lists go through introsort, which permutes equal distances
      if (matches.back().size() > 16) std::sort(matches.back().begin(), matches.back().end());
    }
  }
  void radiusImpl(const agast::Mat& query, const std::vector<agast::Mat>& train, std::vector<std::vector<DMatch> >& matches,
                  float maxDistance, const std::vector<agast::Mat>& masks, bool compactResult) const {
    matches.clear();
    if (query.rows == 0) return;
    Collection c;
    gather(query, train, masks, &c);
    const int nt = c.start.back();
    std::vector<int64_t> offsets((size_t)query.rows + 1, 0);
    std::vector<int32_t> idx, dist;
    if (nt > 0) {
      const std::vector<unsigned char> q = tight(query);
      brisk_ctx* ctx = detail::context();
      int64_t cap = std::max<int64_t>(1024, 4 * (int64_t)query.rows);
      for (;;) {
        idx.resize((size_t)cap); dist.resize((size_t)cap);
        const int rc = brisk_hamming_radius(ctx, q.data(), query.rows, c.all.data(), nt, query.cols, maxDistance,
                                            c.mask.empty() ? nullptr : c.mask.data(), 1, offsets.data(), idx.data(), dist.data(), cap);
        if (rc == BRISK_ERR_CAPACITY && offsets.back() > cap) { cap = offsets.back(); continue; }
        detail::check(ctx, rc);
        break;
      }
    }
    for (int q = 0; q < query.rows; ++q) {
      if (c.masked_out[q]) { if (!compactResult) matches.push_back(std::vector<DMatch>()); continue; }
      matches.push_back(std::vector<DMatch>());
      for (int64_t j = offsets[q]; j < offsets[q + 1]; ++j) matches.back().push_back(make(q, idx[(size_t)j], c.start, (float)dist[(size_t)j]));
    }
  }
#ifndef BRISK_B200_USE_OPENCV
  std::vector<agast::Mat> train_;
#endif
};

// deprecated aliases of the reference (brisk/include/brisk/brisk.h:61-65)
typedef BruteForceMatcher BruteForceMatcherSse;
typedef Hamming HammingSse;

}  // namespace brisk

#ifdef BRISK_B200_USE_OPENCV
namespace cv {  // brisk/include/brisk/brisk.h:56-59
typedef brisk::BriskDescriptorExtractor BriskDescriptorExtractor;
typedef brisk::BriskFeatureDetector BriskFeatureDetector;
}  // namespace cv
#endif
#endif  // BRISK_BRISK_H_
