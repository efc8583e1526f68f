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
  // header for row i sharing the data (cv::Mat::row)
  Mat row(int i) const {
    Mat m;
    m.rows = 1; m.cols = cols; m.type_ = type_; m.step = step; m.data = data + step.buf[0] * i; m.owner = owner;
    return m;
  }
  // cv::Mat::setTo(Scalar) for the single-channel types the reference uses it on (CV_32S, CV_8U)
  template <typename S> Mat& setTo(const S& s) {
    for (int r = 0; r < rows; ++r)
      for (int c = 0; c < cols; ++c) {
        if ((type_ & 7) == CV_32S) at<int>(r, c) = (int)s.val[0];
        else if ((type_ & 7) == CV_32F) at<float>(r, c) = (float)s.val[0];
        else if (esz(type_) == 1) at<unsigned char>(r, c) = (unsigned char)s.val[0];
        else at<unsigned short>(r, c) = (unsigned short)s.val[0];
      }
    return *this;
  }
};

struct Scalar {
  double val[4];
  Scalar() { val[0] = val[1] = val[2] = val[3] = 0; }
  static Scalar all(double v) { Scalar s; s.val[0] = s.val[1] = s.val[2] = s.val[3] = v; return s; }
};

template <typename T> struct DataType;
template <> struct DataType<unsigned char> { enum { type = CV_8UC1 }; };
template <> struct DataType<int> { enum { type = CV_32SC1 }; };
template <> struct DataType<float> { enum { type = CV_32FC1 }; };

// cv::Ptr: the reference only constructs it from a raw pointer and hands it back
template <typename T>
struct Ptr : std::shared_ptr<T> {
  Ptr() {}
  template <typename Y> Ptr(Y* p) : std::shared_ptr<T>(p) {}
};

inline int countNonZero(const Mat& m) {
  int n = 0;
  for (int r = 0; r < m.rows; ++r)
    for (int c = 0; c < m.cols; ++c) n += m.at<unsigned char>(r, c) != 0;
  return n;
}

// cv::minMaxLoc on a single-channel CV_32S matrix: extreme values and the FIRST location (row-major scan,
// strict comparisons) at which each is met, as OpenCV's minMaxIdx does.
inline void minMaxLoc(const Mat& m, double* minVal, double* maxVal, Point* minLoc, Point* maxLoc) {
  int lo = std::numeric_limits<int>::max(), hi = std::numeric_limits<int>::min();
  Point plo(-1, -1), phi(-1, -1);
  for (int r = 0; r < m.rows; ++r)
    for (int c = 0; c < m.cols; ++c) {
      const int v = m.at<int>(r, c);
      if (plo.x < 0 || v < lo) { lo = v; plo = Point(c, r); }
      if (phi.x < 0 || v > hi) { hi = v; phi = Point(c, r); }
    }
  if (minVal) *minVal = lo;
  if (maxVal) *maxVal = hi;
  if (minLoc) *minLoc = plo;
  if (maxLoc) *maxLoc = phi;
}
#define CV_DbgAssert(expr) ((void)0)

struct _InputArray {
  Mat m;
  std::vector<Mat> vec;
  _InputArray() {}
  _InputArray(const Mat& mm) : m(mm) {}
  _InputArray(const std::vector<Mat>& v) : vec(v) {}
  Mat getMat() const { return m; }
  void getMatVector(std::vector<Mat>& out) const { out = vec; }
  bool empty() const { return m.empty() && vec.empty(); }
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
