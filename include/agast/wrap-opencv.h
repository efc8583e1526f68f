// Minimal image / key-point types for builds without OpenCV.
//
// The reference routes all of its OpenCV use through agast/wrap-opencv.h
// (reference agast/include/agast/wrap-opencv.h:41-99: `agast::Mat`,
// `agast::KeyPoint` and the KeyPointX/Y/... accessors are aliases of the cv::
// types when HAVE_OPENCV, else small stand-ins, :100-329).  This header is the
// stand-in branch for the B200 drop-in: callers that have OpenCV can define
// BRISK_B200_USE_OPENCV before including <brisk/brisk.h> and get the cv:: types
// (cv::KeyPoint is layout compatible with brisk_keypoint).
#ifndef AGAST_WRAP_OPENCV_H_
#define AGAST_WRAP_OPENCV_H_

#ifdef BRISK_B200_USE_OPENCV
#include <opencv2/core/core.hpp>
#include <opencv2/features2d/features2d.hpp>
namespace agast {
using cv::KeyPoint;
using cv::Mat;
}  // namespace agast
#else
#include <stdint.h>
#include <cstring>
#include <memory>
#include <vector>

#ifndef CV_8U
#define CV_8U 0
#define CV_8UC1 0
#endif

namespace agast {
struct Point2f { float x, y; };
struct KeyPoint {  // == cv::KeyPoint field order, 28 bytes
  Point2f pt;
  float size, angle, response;
  int octave, class_id;
  KeyPoint() : pt{0, 0}, size(0), angle(-1), response(0), octave(0), class_id(-1) {}
  KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1)
      : pt{x, y}, size(s), angle(a), response(r), octave(o), class_id(c) {}
};
// Continuous 8-bit matrix with shared ownership (the slice of cv::Mat the BRISK API needs).
struct Mat {
  int rows = 0, cols = 0;
  size_t step = 0;
  unsigned char* data = nullptr;
  std::shared_ptr<unsigned char> owner;
  Mat() {}
  Mat(int r, int c, int /*type*/) { create(r, c, CV_8UC1); }
  Mat(int r, int c, int /*type*/, void* d, size_t s = 0) : rows(r), cols(c), step(s ? s : (size_t)c), data((unsigned char*)d) {}
  void create(int r, int c, int /*type*/) {
    rows = r; cols = c; step = (size_t)c;
    owner.reset(new unsigned char[(size_t)r * c + 64], std::default_delete<unsigned char[]>());
    data = owner.get();
  }
  static Mat zeros(int r, int c, int t) { Mat m(r, c, t); std::memset(m.data, 0, (size_t)r * c); return m; }
  bool empty() const { return data == nullptr || rows * cols == 0; }
  int type() const { return CV_8UC1; }
  bool isContinuous() const { return step == (size_t)cols; }
};
}  // namespace agast
#endif  // BRISK_B200_USE_OPENCV

namespace agast {
inline float& KeyPointX(KeyPoint& k) { return k.pt.x; }
inline const float& KeyPointX(const KeyPoint& k) { return k.pt.x; }
inline float& KeyPointY(KeyPoint& k) { return k.pt.y; }
inline const float& KeyPointY(const KeyPoint& k) { return k.pt.y; }
inline float& KeyPointSize(KeyPoint& k) { return k.size; }
inline float& KeyPointAngle(KeyPoint& k) { return k.angle; }
inline float& KeyPointResponse(KeyPoint& k) { return k.response; }
inline int& KeyPointOctave(KeyPoint& k) { return k.octave; }
}  // namespace agast
#endif  // AGAST_WRAP_OPENCV_H_
