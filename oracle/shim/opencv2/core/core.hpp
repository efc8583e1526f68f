// ORACLE SCAFFOLDING (test infrastructure, not product code).
// Minimal stand-in for the slice of the OpenCV core API that the reference hot
// path touches, so that the UNMODIFIED reference sources under /root/reference
// compile here without OpenCV (which this image does not have).  Rows are
// 64-byte aligned because the reference's SSE loops use aligned loads
// (reference brisk/src/image-down-sampling.cc:300-302).
#pragma once
#include <stdint.h>
#include <unistd.h>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <vector>

#define CV_CN_SHIFT 3
#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_MAKETYPE(depth, cn) (((depth) & 7) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8SC1 CV_MAKETYPE(CV_8S, 1)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_16SC1 CV_MAKETYPE(CV_16S, 1)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)

namespace cv {
template <typename T>
struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T a, T b) : x(a), y(b) {}
};
typedef Point_<float> Point2f;
typedef Point_<int> Point2i;
typedef Point2i Point;

struct KeyPoint {
  Point2f pt;
  float size, angle, response;
  int octave, class_id;
  KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
  KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1)
      : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
};

struct MatStep {
  size_t buf[2];
  MatStep() { buf[0] = buf[1] = 0; }
  operator size_t() const { return buf[0]; }
  size_t& operator[](int i) { return buf[i]; }
  const size_t& operator[](int i) const { return buf[i]; }
};

struct Mat {
  int rows, cols, type_;
  MatStep step;
  unsigned char* data;
  std::shared_ptr<unsigned char> owner;
  static int esz(int t) {
    switch (t & 7) {
      case CV_8U: case CV_8S: return 1;
      case CV_16U: case CV_16S: return 2;
      default: return 4;
    }
  }
  Mat() : rows(0), cols(0), type_(0), data(nullptr) {}
  Mat(int r, int c, int t) { create(r, c, t); }
  Mat(int r, int c, int t, void* d) : rows(r), cols(c), type_(t), data((unsigned char*)d) {
    step.buf[0] = (size_t)c * esz(t);
  }
  void create(int r, int c, int t) {
    rows = r; cols = c; type_ = t;
    step.buf[0] = (size_t)c * esz(t);
    size_t n = (size_t)r * c * esz(t) + 64;
    n = (n + 63) / 64 * 64;
    unsigned char* p = (unsigned char*)aligned_alloc(64, n);
    owner.reset(p, free);
    data = p;
  }
  static Mat zeros(int r, int c, int t) {
    Mat m(r, c, t);
    memset(m.data, 0, (size_t)r * c * esz(t));
    return m;
  }
  Mat clone() const {
    Mat m;
    if (!data) return m;
    m.create(rows, cols, type_);
    memcpy(m.data, data, (size_t)rows * cols * esz(type_));
    return m;
  }
  int type() const { return type_; }
  bool empty() const { return data == nullptr || rows * cols == 0; }
  size_t elemSize() const { return esz(type_); }
  bool isContinuous() const { return true; }
  void release() { owner.reset(); data = nullptr; rows = cols = 0; }
  template <typename T> T& at(int i, int j) { return ((T*)(data + step.buf[0] * i))[j]; }
  template <typename T> const T& at(int i, int j) const { return ((const T*)(data + step.buf[0] * i))[j]; }
  template <typename T> T& at(int i) { return ((T*)data)[i]; }
  template <typename T> const T& at(int i) const { return ((const T*)data)[i]; }
};

struct _InputArray {
  Mat m;
  _InputArray() {}
  _InputArray(const Mat& mm) : m(mm) {}
  Mat getMat() const { return m; }
};
struct _OutputArray {
  Mat* p;
  _OutputArray() : p(nullptr) {}
  _OutputArray(Mat& mm) : p(&mm) {}
  Mat& getMatRef() const { return *p; }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
typedef InputArray InputArrayOfArrays;
inline const _InputArray& noArray() { static _InputArray a; return a; }
inline const _OutputArray& noOutArray() { static _OutputArray a; return a; }
enum { IMREAD_GRAYSCALE = 0 };
Mat imread(const std::string& fn, int flags = 0);
}  // namespace cv
