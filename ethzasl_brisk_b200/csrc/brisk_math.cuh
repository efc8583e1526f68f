// Per-pixel / per-corner arithmetic of the BRISK hot path, written as
// __host__ __device__ functions so that the exact code the kernels run can also
// be compiled for the host by the unit tests (tests/host_emul).  All float
// code must be compiled without FMA contraction (nvcc -fmad=false,
// g++ -ffp-contract=off): the reference is built without FMA and its results
// are compared bit for bit.
#pragma once
#include "brisk_common.cuh"

namespace briskb200 {

// ---------------------------------------------------------------------------
// Down-sampling (reference brisk/src/image-down-sampling.cc).
// ---------------------------------------------------------------------------

BRISK_HD int avg_up(int a, int b) { return (a + b + 1) >> 1; }

// One output pixel of Halfsample8 (image-down-sampling.cc:142-392).  `c` is the
// output column and `src_w` the source width: the reference's SSE body rounds
// up twice, an 8-pixel half block truncates horizontally, and the scalar tail
// rounds the 4-sum once.
BRISK_HD int halfsample_px(int a0, int a1, int b0, int b1, int c, int src_w) {
  const int hsize = src_w >> 4;
  const int body = (hsize >> 1) << 4;
  if (c < body) return avg_up(avg_up(a0, b0), avg_up(a1, b1));
  if ((hsize & 1) && c < body + 8) return (avg_up(a0, b0) + avg_up(a1, b1)) >> 1;
  return (a0 + a1 + b0 + b1 + 2) >> 2;
}

// 2x2 outputs of one 3x3 source block of Twothirdsample8
// (image-down-sampling.cc:550-787).  `T` is the triple index along x; triples
// inside the 15-column SSE blocks use average-of-averages, the scalar tail the
// 4:2:2:1 /9 weights.  p = 3x3 block, row-major.
BRISK_HD void twothird_block(const int p[9], int T, int src_w, int out[4]) {
  if (T < 5 * (src_w / 15)) {
    const int u0 = avg_up(avg_up(p[0], p[3]), p[0]), u1 = avg_up(avg_up(p[1], p[4]), p[1]), u2 = avg_up(avg_up(p[2], p[5]), p[2]);
    const int l0 = avg_up(avg_up(p[6], p[3]), p[6]), l1 = avg_up(avg_up(p[7], p[4]), p[7]), l2 = avg_up(avg_up(p[8], p[5]), p[8]);
    out[0] = avg_up(avg_up(u0, u1), u0);
    out[1] = avg_up(avg_up(u2, u1), u2);
    out[2] = avg_up(avg_up(l0, l1), l0);
    out[3] = avg_up(avg_up(l2, l1), l2);
  } else {
    out[0] = (4 * p[0] + 2 * (p[1] + p[3] + 1) + p[4] + 1) / 9;
    out[1] = (4 * p[2] + 2 * (p[1] + p[5] + 1) + p[4] + 1) / 9;
    out[2] = (4 * p[6] + 2 * (p[7] + p[3] + 1) + p[4] + 1) / 9;
    out[3] = (4 * p[8] + 2 * (p[7] + p[5] + 1) + p[4] + 1) / 9;
  }
}

// ---------------------------------------------------------------------------
// FAST / AGAST scores in closed form.
// ---------------------------------------------------------------------------

BRISK_HD int imin(int a, int b) { return a < b ? a : b; }
BRISK_HD int imax(int a, int b) { return a > b ? a : b; }

// m = max over all arcs of ARC contiguous ring pixels of
// max(min_i(p_i - c), min_i(c - p_i)).  The generated AGAST trees
// (agast/src/oast9-16.cc:43-1859) and the bisection cornerScore
// (oast9-16-nms.cc:39-1976, agast5-8-nms.cc:39-358) reduce to this quantity:
// is-corner(b) <=> m-1 >= b, cornerScore(b) = max(b, m-1).
//
// Written on the raw pixel values as max(min(p) - c, c - max(p)).  Do NOT
// rewrite it as max(min(d), -max(d)) on differences: ptxas 12.9 for sm_100a
// folds that negation into VIMNMX3 and drops it (observed on hardware: wrong
// scores), while plain subtractions are compiled correctly.
BRISK_HD int imin3(int a, int b, int c) { return imin(imin(a, b), c); }  // one VIMNMX3 on sm_100a
BRISK_HD int imax3(int a, int b, int c) { return imax(imax(a, b), c); }

BRISK_HD int arc_contrast16(const int p[16], int c) {
  // sliding minimum / maximum over windows of 9 on the circular sequence, as 3-input operations:
  // windows of 3, then three of those 3 apart; max_i(min9_i) - c and c - min_i(max9_i) are then
  // the best bright / dark arc contrasts
  int lo3[16], hi3[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { lo3[i] = imin3(p[i], p[(i + 1) & 15], p[(i + 2) & 15]); hi3[i] = imax3(p[i], p[(i + 1) & 15], p[(i + 2) & 15]); }
  int lo9[16], hi9[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { lo9[i] = imin3(lo3[i], lo3[(i + 3) & 15], lo3[(i + 6) & 15]); hi9[i] = imax3(hi3[i], hi3[(i + 3) & 15], hi3[(i + 6) & 15]); }
  int bright = lo9[15], dark = hi9[15];
#pragma unroll
  for (int i = 0; i < 15; i += 3) { bright = imax(bright, imax3(lo9[i], lo9[i + 1], lo9[i + 2])); dark = imin(dark, imin3(hi9[i], hi9[i + 1], hi9[i + 2])); }
  return imax(bright - c, c - dark);
}

BRISK_HD int arc_contrast8(const int p[8], int c) {
  int best = -1000;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int lo = p[i], hi = p[i];
#pragma unroll
    for (int k = 1; k < 5; ++k) { lo = imin(lo, p[(i + k) & 7]); hi = imax(hi, p[(i + k) & 7]); }
    best = imax(best, imax(lo - c, c - hi));
  }
  return best;
}

// FAST 9-16 score F = m-1 at (x, y); ring of agast/include/agast/oast9-16.h:99-116.
// The caller guarantees a 3-pixel margin.
BRISK_HD int fast916(const uint8_t* img, int pitch, int x, int y) {
  const uint8_t* p = img + (long long)y * pitch + x;
  int d[16];
  d[0] = p[-3];              d[1] = p[-3 - pitch];      d[2] = p[-2 - 2 * pitch];  d[3] = p[-1 - 3 * pitch];
  d[4] = p[-3 * pitch];      d[5] = p[1 - 3 * pitch];   d[6] = p[2 - 2 * pitch];   d[7] = p[3 - pitch];
  d[8] = p[3];               d[9] = p[3 + pitch];       d[10] = p[2 + 2 * pitch];  d[11] = p[1 + 3 * pitch];
  d[12] = p[3 * pitch];      d[13] = p[-1 + 3 * pitch]; d[14] = p[-2 + 2 * pitch]; d[15] = p[-3 + pitch];
  return arc_contrast16(d, p[0]) - 1;
}

// AGAST 5-8 score (ring of agast/include/agast/agast5-8.h:68-77); 1-pixel margin.
BRISK_HD int fast58(const uint8_t* img, int pitch, int x, int y) {
  const uint8_t* p = img + (long long)y * pitch + x;
  int d[8];
  d[0] = p[-1];         d[1] = p[-1 - pitch]; d[2] = p[-pitch];     d[3] = p[1 - pitch];
  d[4] = p[1];          d[5] = p[1 + pitch];  d[6] = p[pitch];      d[7] = p[-1 + pitch];
  return arc_contrast8(d, p[0]) - 1;
}

// Segment test of OastDetector9_16::detect with the per-pixel adaptive threshold
// (agast/src/oast9-16.cc:86-100, ast-detector.h:62-68): T = threshold-map value,
// b = user threshold.  Returns true when (x, y) is a corner.
BRISK_HD bool agast_is_corner(const uint8_t* img, int pitch, int x, int y, int T, int b, int lower = kLowerThreshold) {
  if (T < (b * lower) / 100) return false;
  const int t = T < lower ? lower : (T > kUpperThreshold ? kUpperThreshold : T);
  const int b2 = (t * b) / 100;
  return fast916(img, pitch, x, y) >= b2;
}

// Threshold map value at (x, y), 3 <= x < w-3, 3 <= y < h-3 (reference
// brisk-layer.cc:278-598): max - min over the centre, the
// four (+-2,+-2) diagonals and the 3x3 blocks centred at (x,y+-2), (x+-2,y).
// Scalar form used by the host emulation; the detect kernel computes the same
// quantity from separable 3x3 min/max planes in shared memory.
BRISK_HD int thrmap_px(const uint8_t* img, int pitch, int x, int y) {
  const uint8_t* p = img + (long long)y * pitch + x;
  int hi = p[0], lo = p[0];
#define BRISK_MM(v) { const int vv = (v); hi = imax(hi, vv); lo = imin(lo, vv); }
  BRISK_MM(p[-2 - 2 * pitch]) BRISK_MM(p[2 - 2 * pitch]) BRISK_MM(p[2 + 2 * pitch]) BRISK_MM(p[-2 + 2 * pitch])
  for (int k = 0; k < 4; ++k) {
    const int cx = (k == 2) ? -2 : (k == 3 ? 2 : 0), cy = (k == 0) ? -2 : (k == 1 ? 2 : 0);
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) BRISK_MM(p[(cy + dy) * pitch + cx + dx])
  }
#undef BRISK_MM
  return hi - lo;
}

// ---------------------------------------------------------------------------
// Sub-pixel / scale refinement (reference brisk/src/brisk-scale-space.cc).
// ---------------------------------------------------------------------------

// brisk-scale-space.cc:1230-1364 (Subpixel2D on int scores).  Argument names
// follow the reference: sXY with X = column index, Y = row index of the 3x3
// patch (s01 is the left neighbour).  Reproduces the `delta_y = delta_x`
// assignment of the boundary case.
BRISK_HD float subpixel2d(int s00, int s01, int s02, int s10, int s11, int s12, int s20, int s21, int s22, float* dx,
                          float* dy) {
  const int t1 = s00 + s02 - 2 * s11 + s20 + s22;
  const int c1 = 3 * (t1 + s01 - ((s10 + s12) << 1) + s21);
  const int c2 = 3 * (t1 - ((s01 + s21) << 1) + s10 + s12);
  const int t2 = s02 - s20;
  const int t3 = s00 + t2 - s22;
  const int t4 = t3 - 2 * t2;
  const int c3 = -3 * (t3 + s01 - s21);
  const int c4 = -3 * (t4 + s10 - s12);
  const int c5 = (s00 - s02 - s20 + s22) * 4;
  const int c6 = (-(s00 + s02 - ((s10 + s01 + s12 + s21) << 1) - 5 * s11 + s20 + s22)) * 2;
  const int H = 4 * c1 * c2 - c5 * c5;
  if (H == 0) {
    *dx = 0.0f; *dy = 0.0f;
    return (float)((double)(float)c6 / 18.0);
  }
  if (!(H > 0 && c1 < 0)) {
    int best = c3 + c4 + c5;
    float bx = 1.0f, by = 1.0f;
    int t = -c3 + c4 - c5;
    if (t > best) { best = t; bx = -1.0f; by = 1.0f; }
    t = c3 - c4 - c5;
    if (t > best) { best = t; bx = 1.0f; by = -1.0f; }
    t = -c3 - c4 + c5;
    if (t > best) { best = t; bx = -1.0f; by = -1.0f; }
    *dx = bx; *dy = by;
    return (float)((double)(float)(best + c1 + c2 + c6) / 18.0);
  }
  const float ddx = (float)(2 * c2 * c3 - c4 * c5) / (float)(-H);
  const float ddy = (float)(2 * c1 * c4 - c3 * c5) / (float)(-H);
  const bool tx = ddx > 1.0f, tx_ = !tx && ddx < -1.0f, ty = ddy > 1.0f, ty_ = ddy < -1.0f;
#define BRISK_QUAD(ax, ay)                                                                                     \
  ((float)((double)((((((float)c1 * (ax)) * (ax) + ((float)c2 * (ay)) * (ay)) + (float)c3 * (ax)) + (float)c4 * (ay)) + \
                     ((float)c5 * (ax)) * (ay) + (float)c6) / 18.0))
  if (tx || tx_ || ty || ty_) {
    float x1 = 0.0f, x2 = 0.0f, y1 = 0.0f, y2 = 0.0f;
    if (tx) {
      x1 = 1.0f; y1 = -(float)(c4 + c5) / (float)(2 * c2);
      y1 = y1 > 1.0f ? 1.0f : (y1 < -1.0f ? -1.0f : y1);
    } else if (tx_) {
      x1 = -1.0f; y1 = -(float)(c4 - c5) / (float)(2 * c2);
      y1 = y1 > 1.0f ? 1.0f : (y1 < -1.0f ? -1.0f : y1);
    }
    if (ty) {
      y2 = 1.0f; x2 = -(float)(c3 + c5) / (float)(2 * c1);
      x2 = x2 > 1.0f ? 1.0f : (x2 < -1.0f ? -1.0f : x2);
    } else if (ty_) {
      y2 = -1.0f; x2 = -(float)(c3 - c5) / (float)(2 * c1);
      x2 = x2 > 1.0f ? 1.0f : (x2 < -1.0f ? -1.0f : x2);
    }
    const float m1 = BRISK_QUAD(x1, y1), m2 = BRISK_QUAD(x2, y2);
    if (m1 > m2) { *dx = x1; *dy = x1; return m1; }
    *dx = x2; *dy = x2;
    return m2;
  }
  *dx = ddx; *dy = ddy;
  return BRISK_QUAD(ddx, ddy);
#undef BRISK_QUAD
}

// brisk-scale-space.cc:1101-1228: parabola through the scores at three scales.
// kind 0 = Refine1D (0.75, 1, 1.5; around an octave), 1 = Refine1D_1 (2/3, 1,
// 4/3; around an intra-octave), 2 = Refine1D_2 (0.7, 1, 1.5; octave 0).
BRISK_HD float refine1d(int kind, float s_05, float s0, float s05, float* max) {
  const int i_05 = (int)(1024.0 * (double)s_05 + 0.5), i0 = (int)(1024.0 * (double)s0 + 0.5), i05 = (int)(1024.0 * (double)s05 + 0.5);
  int a, b, c;
  double lo, hi;
  if (kind == 0) {
    a = 16 * i_05 - 24 * i0 + 8 * i05; b = -40 * i_05 + 54 * i0 - 14 * i05; c = 24 * i_05 - 27 * i0 + 6 * i05;
    lo = 0.75; hi = 1.5;
  } else if (kind == 1) {
    a = 9 * i_05 - 18 * i0 + 9 * i05; b = -21 * i_05 + 36 * i0 - 15 * i05; c = 12 * i_05 - 16 * i0 + 6 * i05;
    lo = 0.6666666666666666666666666667; hi = 1.3333333333333333333333333333;
  } else {
    a = 2 * i_05 - 4 * i0 + 2 * i05; b = -5 * i_05 + 8 * i0 - 3 * i05; c = 3 * i_05 - 3 * i0 + 1 * i05;
    lo = 0.7; hi = 1.5;
  }
  if (a >= 0) {
    if (s0 >= s_05 && s0 >= s05) { *max = s0; return 1.0f; }
    if (s_05 >= s0 && s_05 >= s05) { *max = s_05; return (float)lo; }
    if (s05 >= s0 && s05 >= s_05) { *max = s05; return (float)hi; }
  }
  float r = -(float)b / (float)(2 * a);
  if ((double)r < lo) r = (float)lo;
  else if ((double)r > hi) r = (float)hi;
  float m = ((float)c + ((float)a * r) * r) + (float)b * r;
  if (kind == 0) m = (float)((double)m / 3072.0);
  else if (kind == 1) m = (float)((double)m / 2048.0);
  else m = m / 1024.0f;
  *max = m;
  return r;
}

// Eight descriptor bits -> eight E2M1 values in one 32-bit word, bit i of the byte in nibble i: +1.0 (0x2) for a set bit,
// -1.0 (0xA) for a clear one (the operands of the FP4 matcher, hamming_tc5.cu).
BRISK_HD uint32_t e2m1_expand_byte(uint32_t byte) {
  // bit i -> bit 4 i: spread the byte's bits four apart
  uint32_t sp = (byte | (byte << 12)) & 0x000f000fu;   // bits 0..3 | bits 4..7 at 16
  sp = (sp | (sp << 6)) & 0x03030303u;                 // two bits per byte
  sp = (sp | (sp << 3)) & 0x11111111u;                 // one bit per nibble
  return 0x22222222u | ((~sp & 0x11111111u) << 3);     // set -> 0x2, clear -> 0xA (the sign bit of the nibble)
}

}  // namespace briskb200
