// Harris scale-space detector arithmetic (reference
// brisk/src/harris-scores.cc:53-279, harris-score-calculator.{h,cc},
// brisk/include/brisk/internal/scale-space-layer-inl.h:193-693,
// uniformity-enforcement-inl.h:44-194) as __host__ __device__ functions shared by
// the kernels in harris.cu and the host-side tests.
#pragma once
#include <math.h>

#include "brisk_common.cuh"
#include "gcc_sort.cuh"

namespace briskb200 {

// {score, x, y} record of the reference (score-calculator.h:66-85, 8 bytes).
struct HPoint {
  int score;
  unsigned short x, y;
};

// Scharr-like derivatives of HarrisScoresSSE (harris-scores.cc:76-160, SURVEY.md
// App. A.5) at one pixel, already shifted left by 3, as 16-bit values; then the
// high halves of their 16x16 products (pmulhw).  p = 3x3 neighbourhood, row-major.
BRISK_HD void harris_products(const int p[9], int* xx, int* yy, int* xy) {
  const int dx = (10 * (p[3] - p[5]) + 3 * (p[0] - p[2]) + 3 * (p[6] - p[8])) * 8;
  const int dy = (10 * (p[1] - p[7]) + 3 * (p[0] - p[6]) + 3 * (p[2] - p[8])) * 8;
  const int sdx = (short)dx, sdy = (short)dy;
  *xx = (sdx * sdx) >> 16;
  *yy = (sdy * sdy) >> 16;
  *xy = (sdx * sdy) >> 16;
}

// 3x3 binomial smoothing with arithmetic >> 4 (harris-scores.cc:162-250); q = 3x3 values.
BRISK_HD int harris_smooth(const int q[9]) {
  return (4 * q[4] + 2 * (q[1] + q[7] + q[3] + q[5]) + q[0] + q[2] + q[6] + q[8]) >> 4;
}

// det - trace^2/16 in the reference's integer form (harris-scores.cc:252-274).
BRISK_HD int harris_response(int a, int b, int c) {
  const int t = (a + b) >> 1;
  return a * b - c * c - ((t * t) >> 2);
}

// HarrisScoreCalculator::Score(double, double) (harris-score-calculator.h:57-74):
// bilinear read of the int32 score map in double, 0 outside.
BRISK_HD double harris_score_bilinear(const int* scores, int pitch, int w, int h, double u, double v) {
  const int ui = (int)u, vi = (int)v;
  if (ui + 1 >= w || vi + 1 >= h || ui < 0 || vi < 0) return 0.0;
  const double ru = u - (double)ui, rv = v - (double)vi;
  const double mu = 1.0 - ru, mv = 1.0 - rv;
  const int* p = scores + (long long)vi * pitch + ui;
  return mv * (mu * (double)p[0] + ru * (double)p[1]) + rv * (mu * (double)p[pitch] + ru * (double)p[pitch + 1]);
}

// x86 cvttsd2si: out-of-range / NaN doubles convert to INT_MIN (the reference
// relies on it through plain `int tmp = double_expr`).
BRISK_HD int d2i_x86(double v) {
  if (!(v > -2147483649.0 && v < 2147483648.0)) return -2147483647 - 1;
  return (int)v;
}

// Subpixel2D on doubles (scale-space-layer-inl.h:559-693); only delta_x / delta_y
// are used by the caller.  Argument order as the reference's call (:395-403):
// rows top to bottom, each left to right.
BRISK_HD void harris_subpixel2d(double s00, double s01, double s02, double s10, double s11, double s12, double s20,
                                double s21, double s22, float* dx, float* dy) {
  const double t1 = s00 + s02 - 2 * s11 + s20 + s22;
  const double c1 = 3 * (t1 + s01 - ((s10 + s12) / 2.0) + s21);
  const double c2 = 3 * (t1 - ((s01 + s21) / 2.0) + s10 + s12);
  const double t2 = s02 - s20;
  const double t3 = s00 + t2 - s22;
  const double t4 = t3 - 2 * t2;
  const double c3 = -3 * (t3 + s01 - s21);
  const double c4 = -3 * (t4 + s10 - s12);
  const double c5 = (s00 - s02 - s20 + s22) / 4.0;
  const double c6 = -(s00 + s02 - ((s10 + s01 + s12 + s21) / 2.0) - 5 * s11 + s20 + s22) / 2.01;
  const double H = 4 * c1 * c2 - c5 * c5;
  if (H == 0) { *dx = 0.0f; *dy = 0.0f; return; }
  if (!(H > 0 && c1 < 0)) {
    int best = d2i_x86(c3 + c4 + c5);
    float bx = 1.0f, by = 1.0f;
    int t = d2i_x86(-c3 + c4 - c5);
    if (t > best) { best = t; bx = -1.0f; by = 1.0f; }
    t = d2i_x86(c3 - c4 - c5);
    if (t > best) { best = t; bx = 1.0f; by = -1.0f; }
    t = d2i_x86(-c3 - c4 + c5);
    if (t > best) { best = t; bx = -1.0f; by = -1.0f; }
    *dx = bx; *dy = by;
    return;
  }
  const float ddx = (float)(2 * c2 * c3 - c4 * c5) / (float)(-H);
  const float ddy = (float)(2 * c1 * c4 - c3 * c5) / (float)(-H);
  const bool tx = ddx > 1.0f, tx_ = !tx && ddx < -1.0f, ty = ddy > 1.0f, ty_ = ddy < -1.0f;
  if (tx || tx_ || ty || ty_) {
    float x1 = 0.0f, x2 = 0.0f, y1 = 0.0f, y2 = 0.0f;
    if (tx) { x1 = 1.0f; y1 = -(float)(c4 + c5) / (float)(2 * c2); y1 = y1 > 1.0f ? 1.0f : (y1 < -1.0f ? -1.0f : y1); }
    else if (tx_) { x1 = -1.0f; y1 = -(float)(c4 - c5) / (float)(2 * c2); y1 = y1 > 1.0f ? 1.0f : (y1 < -1.0f ? -1.0f : y1); }
    if (ty) { y2 = 1.0f; x2 = -(float)(c3 + c5) / (float)(2 * c1); x2 = x2 > 1.0f ? 1.0f : (x2 < -1.0f ? -1.0f : x2); }
    else if (ty_) { y2 = -1.0f; x2 = -(float)(c3 - c5) / (float)(2 * c1); x2 = x2 > 1.0f ? 1.0f : (x2 < -1.0f ? -1.0f : x2); }
    // double arithmetic with float deltas promoted, left to right, result rounded to float
    const float m1 = (float)((((((c1 * (double)x1) * (double)x1 + (c2 * (double)y1) * (double)y1) + c3 * (double)x1) + c4 * (double)y1) +
                              (c5 * (double)x1) * (double)y1 + c6) / 18.0);
    const float m2 = (float)((((((c1 * (double)x2) * (double)x2 + (c2 * (double)y2) * (double)y2) + c3 * (double)x2) + c4 * (double)y2) +
                              (c5 * (double)x2) * (double)y2 + c6) / 18.0);
    if (m1 > m2) { *dx = x1; *dy = x1; } else { *dx = x2; *dy = x2; }
    return;
  }
  *dx = ddx; *dy = ddy;
}

// ---------------------------------------------------------------------------
// std::sort on HPoint with the reference's comparator `a.score > b.score`
// (score-calculator.h:82-84): see gcc_sort.cuh.  Equal Harris scores are the norm, so
// the exact permutation matters: it fixes both the greedy uniformity outcome and the output order.
// ---------------------------------------------------------------------------
struct HpLess { BRISK_HD bool operator()(const HPoint& a, const HPoint& b) const { return a.score > b.score; } };
BRISK_HD bool hp_before(const HPoint& a, const HPoint& b) { return HpLess()(a, b); }
BRISK_HD void hp_heapsort(HPoint* first, long len) { gs_heapsort(HpLess(), first, len); }
BRISK_HD void hp_insertion_sort(HPoint* first, HPoint* last) { gs_insertion_sort(HpLess(), first, last); }
BRISK_HD int gcc_partition(HPoint* a, int first, int last) { return gs_partition(HpLess(), a, first, last); }
BRISK_HD void gcc_sort_range(HPoint* a, int first0, int last0, int depth0) { gs_sort_range(HpLess(), a, first0, last0, depth0); }
BRISK_HD int gcc_depth_limit(int n) { return gs_depth_limit(n); }
BRISK_HD void gcc_sort(HPoint* a, int n) { gs_sort(HpLess(), a, n); }

// Uniformity enforcement constants (uniformity-enforcement-inl.h:59-80).
// LUT(y,x) = max(1 - ((15-x)^2 + (15-y)^2) / 225, 0) as float (scale-space-layer-inl.h:89-97).
BRISK_HD float uniformity_lut(int x, int y) {
  const double v = 1 - (double)((15 - x) * (15 - x) + (15 - y) * (15 - y)) / (double)(15 * 15);
  return (float)(v > 0.0 ? v : 0.0);
}
// normalised score of a point (:79): sqrtf(sqrtf(score / maxScore)) * 255
BRISK_HD float uniformity_nsc1(int score, float max_score) { return sqrtf(sqrtf((float)score / max_score)) * 255.0f; }
// occupancy increment of LUT cell (x, y) for a point with nsc = 0.99f * nsc1 (:88-170)
BRISK_HD int uniformity_stamp(int x, int y, float nsc) { return (int)ceilf(uniformity_lut(x, y) * nsc) & 0xff; }

}  // namespace briskb200
