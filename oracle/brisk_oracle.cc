// ORACLE (TEST INFRASTRUCTURE ONLY -- never linked into or called by the product).
//
// CPU restatement of the reference's BRISK hot path (ethz-asl/ethzasl_brisk) in
// plain closed-form scalar C++.  It follows the reference's *results*, including
// its per-column rounding regimes, its order-dependent lazy score cache and its
// known bugs, but shares no code with it: the generated
// AGAST decision trees become min/max arc arithmetic, the SSE loops become
// per-pixel formulas.  Each function cites the reference file:line it restates.
//
// Pinning: tests/test_oracle_golden.py checks this file against the reference's
// own golden fixtures (brisk_verification_{ast,harris}.set, committed as
// tests/golden/brisk_verification.npz) and, in the same file (`*_vs_ref` tests) and in the seeded random sweeps of
// tests/test_host_logic.py, stage by stage against the unmodified reference compiled into oracle/_ref.
//
// NB: this file includes <math.h>, which in C++ also exposes the float
// overloads of log/sqrt/atan2 in the global namespace; the reference's
// unqualified calls bind to the DOUBLE C functions (its objects import only
// log, pow, sin, sincos, atan2, sqrt and sqrtf), so every such call below casts
// its argument to double explicitly.
//
// Build: oracle/Makefile (`make oracle`), -ffp-contract=off, SSE2 scalar floats
// only (the reference is built without FMA).

#include "brisk_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <memory>
#include <mutex>
#include <vector>

#include "brisk_pattern_data.inc"

namespace {

typedef std::vector<uint8_t> Bytes;

// ---------------------------------------------------------------------------
// Down-sampling.
// ---------------------------------------------------------------------------

inline int Avg(int a, int b) { return (a + b + 1) >> 1; }  // pavgb

// reference image-down-sampling.cc:142-392 (Halfsample8).  Three rounding
// regimes by output column:
//   [0, 16*(W/32))            double round-up:   avg(avg(a0,b0), avg(a1,b1))
//   next 8 if (W/16) is odd   vertical round-up, horizontal truncation
//   last (W%16)/2             (a0+a1+b0+b1+2)/4
void Halfsample8(const uint8_t* src, int w, int h, uint8_t* dst) {
  const int dw = w / 2, dh = h / 2;
  const int hsize = w / 16;
  const int body = 16 * (hsize / 2);
  const int half_end = (hsize % 2) ? 8 : 0;
  for (int r = 0; r < dh; ++r) {
    const uint8_t* a = src + (size_t)(2 * r) * w;
    const uint8_t* b = a + w;
    uint8_t* d = dst + (size_t)r * dw;
    for (int c = 0; c < dw; ++c) {
      const int a0 = a[2 * c], a1 = a[2 * c + 1], b0 = b[2 * c], b1 = b[2 * c + 1];
      int v;
      if (c < body) v = Avg(Avg(a0, b0), Avg(a1, b1));
      else if (c < body + half_end) v = (Avg(a0, b0) + Avg(a1, b1)) / 2;
      else v = (a0 + a1 + b0 + b1 + 2) / 4;
      d[c] = (uint8_t)v;
    }
  }
}

// reference image-down-sampling.cc:550-787 (Twothirdsample8).  Source triples
// inside the 15-column SSE blocks use 3:1 average-of-averages, the rest the
// 4:2:2:1 /9 integer weights.
void Twothirdsample8(const uint8_t* src, int w, int h, uint8_t* dst) {
  const int dw = 2 * (w / 3);
  const int sse_triples = 5 * (w / 15);
  for (int R = 0; R < h / 3; ++R) {
    const uint8_t* r0 = src + (size_t)(3 * R) * w;
    const uint8_t* r1 = r0 + w;
    const uint8_t* r2 = r1 + w;
    uint8_t* d0 = dst + (size_t)(2 * R) * dw;
    uint8_t* d1 = d0 + dw;
    for (int T = 0; T < w / 3; ++T) {
      const int A1 = r0[3 * T], A2 = r0[3 * T + 1], A3 = r0[3 * T + 2];
      const int B1 = r1[3 * T], B2 = r1[3 * T + 1], B3 = r1[3 * T + 2];
      const int C1 = r2[3 * T], C2 = r2[3 * T + 1], C3 = r2[3 * T + 2];
      if (T < sse_triples) {
        const int u0 = Avg(Avg(A1, B1), A1), u1 = Avg(Avg(A2, B2), A2), u2 = Avg(Avg(A3, B3), A3);
        const int l0 = Avg(Avg(C1, B1), C1), l1 = Avg(Avg(C2, B2), C2), l2 = Avg(Avg(C3, B3), C3);
        d0[2 * T] = (uint8_t)Avg(Avg(u0, u1), u0);
        d0[2 * T + 1] = (uint8_t)Avg(Avg(u2, u1), u2);
        d1[2 * T] = (uint8_t)Avg(Avg(l0, l1), l0);
        d1[2 * T + 1] = (uint8_t)Avg(Avg(l2, l1), l2);
      } else {
        d0[2 * T] = (uint8_t)((4 * A1 + 2 * (A2 + B1 + 1) + B2 + 1) / 9);
        d0[2 * T + 1] = (uint8_t)((4 * A3 + 2 * (A2 + B3 + 1) + B2 + 1) / 9);
        d1[2 * T] = (uint8_t)((4 * C1 + 2 * (C2 + B1 + 1) + B2 + 1) / 9);
        d1[2 * T + 1] = (uint8_t)((4 * C3 + 2 * (C2 + B3 + 1) + B2 + 1) / 9);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Threshold map, FAST scores, corner detection.
// ---------------------------------------------------------------------------

// reference brisk-layer.cc:278-598 (CalculateThresholdMap), SURVEY.md App. A.3:
// local contrast = max - min over the centre, the four (+-2,+-2) diagonals and
// the 3x3 blocks centred at (x, y+-2) and (x+-2, y); zero in the 3-pixel border.
void ThresholdMap(const uint8_t* img, int w, int h, uint8_t* thr) {
  memset(thr, 0, (size_t)w * h);
  if (w < 7 || h < 7) return;
  Bytes mx((size_t)w * h, 0), mn((size_t)w * h, 0);
  for (int y = 1; y < h - 1; ++y)
    for (int x = 1; x < w - 1; ++x) {
      int hi = 0, lo = 255;
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          const int v = img[(size_t)(y + dy) * w + x + dx];
          hi = std::max(hi, v); lo = std::min(lo, v);
        }
      mx[(size_t)y * w + x] = (uint8_t)hi; mn[(size_t)y * w + x] = (uint8_t)lo;
    }
  for (int y = 3; y < h - 3; ++y)
    for (int x = 3; x < w - 3; ++x) {
      int hi = img[(size_t)y * w + x], lo = hi;
      static const int kDiag[4][2] = {{-2, -2}, {2, -2}, {2, 2}, {-2, 2}};
      for (auto& d : kDiag) {
        const int v = img[(size_t)(y + d[1]) * w + x + d[0]];
        hi = std::max(hi, v); lo = std::min(lo, v);
      }
      static const int kAxial[4][2] = {{0, -2}, {0, 2}, {-2, 0}, {2, 0}};
      for (auto& d : kAxial) {
        const size_t o = (size_t)(y + d[1]) * w + x + d[0];
        hi = std::max<int>(hi, mx[o]); lo = std::min<int>(lo, mn[o]);
      }
      thr[(size_t)y * w + x] = (uint8_t)(hi - lo);
    }
}

// Bresenham circle of radius 3, reference agast/include/agast/oast9-16.h:99-116.
const int kRing16[16][2] = {{-3, 0}, {-3, -1}, {-2, -2}, {-1, -3}, {0, -3}, {1, -3}, {2, -2}, {3, -1},
                            {3, 0},  {3, 1},   {2, 2},   {1, 3},   {0, 3},  {-1, 3}, {-2, 2}, {-3, 1}};
// 8-neighbourhood ring, reference agast/include/agast/agast5-8.h:68-77.
const int kRing8[8][2] = {{-1, 0}, {-1, -1}, {0, -1}, {1, -1}, {1, 0}, {1, 1}, {0, 1}, {-1, 1}};

// Largest m such that some arc of `arc` contiguous ring pixels is entirely
// brighter than centre+m-1 ... i.e. m = max over arcs of
// max(min_i(p_i - c), min_i(c - p_i)).  The generated decision trees of
// reference agast/src/oast9-16.cc:43-1859 (arc=9, n=16) and the bisection
// cornerScore of oast9-16-nms.cc:39-1976 / agast5-8-nms.cc:39-358 reduce to
// this: is-corner(b) <=> m - 1 >= b; cornerScore(b) = max(b, m-1).
template <int N, int ARC>
int ArcContrast(const uint8_t* img, int w, int x, int y, const int (*ring)[2]) {
  int d[N];
  const int c = img[(size_t)y * w + x];
  for (int i = 0; i < N; ++i) d[i] = (int)img[(size_t)(y + ring[i][1]) * w + x + ring[i][0]] - c;
  int best = -1000;
  for (int s = 0; s < N; ++s) {
    int lo = 1000, hi = -1000;
    for (int i = 0; i < ARC; ++i) {
      const int v = d[(s + i) % N];
      lo = std::min(lo, v); hi = std::max(hi, v);
    }
    best = std::max(best, std::max(lo, -hi));
  }
  return best;
}

inline int Fast916(const uint8_t* img, int w, int x, int y) { return ArcContrast<16, 9>(img, w, x, y, kRing16) - 1; }
inline int Fast58(const uint8_t* img, int w, int x, int y) { return ArcContrast<8, 5>(img, w, x, y, kRing8) - 1; }

struct Corner { int x, y, score; };

// reference oast9-16.cc:79-100,1844-1856 (detect with threshold map) +
// brisk-layer.cc:99-117 (GetAgastPoints: score at a corner = cornerScore with
// b = thrmap value, which is the thrmap value itself).
void DetectCorners(const uint8_t* img, const uint8_t* thr, int w, int h, int b, int lower, int upper,
                   std::vector<Corner>* out) {
  out->clear();
  const int cmp = (b * lower) / 100;  // ast-detector.h:62-68
  for (int y = 3; y < h - 3; ++y)
    for (int x = 3; x <= w - 4; ++x) {
      int t = thr[(size_t)y * w + x];
      if (t < cmp) continue;
      const int raw = t;
      t = std::min(std::max(t, lower), upper);
      const int b2 = (t * b) / 100;
      const int f = Fast916(img, w, x, y);
      if (f >= b2) out->push_back(Corner{x, y, std::max(raw, f)});
    }
}

// ---------------------------------------------------------------------------
// Scale space (AGAST).
// ---------------------------------------------------------------------------

struct Layer {
  int w = 0, h = 0;
  Bytes img, thr, scores;
  float scale = 1.0f, offset = 0.0f;

  void Init(int b_unused) { (void)b_unused; thr.resize((size_t)w * h); scores.assign((size_t)w * h, 0); ThresholdMap(img.data(), w, h, thr.data()); }

  // reference brisk-layer.cc:118-132: lazy, cached FAST score.  A cached value
  // > 2 is returned whatever threshold is asked.
  int Score(int x, int y, int threshold) {
    if (x < 3 || y < 3 || x >= w - 3 || y >= h - 3) return 0;
    uint8_t& s = scores[(size_t)y * w + x];
    if (s > 2) return s;
    // the byte is assigned first and compared afterwards (:128-130): with threshold 0 (only asked by the
    // "provided key points" mode, brisk-scale-space.cc:119) a pixel that is no corner at b = 0 scores
    // cornerScore = -1, is stored as 255 and stays cached
    const int v = std::max(threshold - 1, Fast916(img.data(), w, x, y));
    s = (uint8_t)v;
    if (s < threshold) s = 0;
    return s;
  }
  // reference brisk-layer.cc:134-145: uncached 5-8 score, border 2.
  int Score58(int x, int y, int threshold) const {
    if (x < 2 || y < 2 || x >= w - 2 || y >= h - 2) return 0;
    const int v = std::max(threshold - 1, Fast58(img.data(), w, x, y));
    return (uint8_t)(v < threshold ? 0 : v);
  }
  // reference brisk-layer.cc:147-161: bilinear interpolation of four lazy
  // scores in float, truncated to a byte.
  int ScoreF(float xf, float yf, int threshold) {
    const int x = (int)xf;
    const float rx1 = xf - (float)x;
    const float rx = 1.0f - rx1;
    const int y = (int)yf;
    const float ry1 = yf - (float)y;
    const float ry = 1.0f - ry1;
    const int s00 = Score(x, y, threshold), s10 = Score(x + 1, y, threshold);
    const int s01 = Score(x, y + 1, threshold), s11 = Score(x + 1, y + 1, threshold);
    const float v = rx * ry * s00 + rx1 * ry * s10 + rx * ry1 * s01 + rx1 * ry1 * s11;
    return (uint8_t)v;
  }
};

// reference brisk-scale-space.cc:1230-1364 (Subpixel2D, int inputs), including
// the delta_y = delta_x assignment in the boundary case.
// Arguments are named s[column][row] as in the reference: s01 = (x-1, y).
float Subpixel2D(int s00, int s01, int s02, int s10, int s11, int s12, int s20, int s21, int s22, float* dx,
                 float* dy) {
  const int t1 = s00 + s02 - 2 * s11 + s20 + s22;
  const int c1 = 3 * (t1 + s01 - ((s10 + s12) << 1) + s21);
  const int c2 = 3 * (t1 - ((s01 + s21) << 1) + s10 + s12);
  const int t2 = s02 - s20;
  const int t3 = s00 + t2 - s22;
  const int t4 = t3 - 2 * t2;
  const int c3 = -3 * (t3 + s01 - s21);
  const int c4 = -3 * (t4 + s10 - s12);
  const int c5 = (s00 - s02 - s20 + s22) << 2;
  const int c6 = -(s00 + s02 - ((s10 + s01 + s12 + s21) << 1) - 5 * s11 + s20 + s22) << 1;
  const int H = 4 * c1 * c2 - c5 * c5;
  if (H == 0) { *dx = 0.0f; *dy = 0.0f; return (float)((double)(float)c6 / 18.0); }
  if (!(H > 0 && c1 < 0)) {
    int best = c3 + c4 + c5; *dx = 1.0f; *dy = 1.0f;
    int t = -c3 + c4 - c5; if (t > best) { best = t; *dx = -1.0f; *dy = 1.0f; }
    t = c3 - c4 - c5;      if (t > best) { best = t; *dx = 1.0f; *dy = -1.0f; }
    t = -c3 - c4 + c5;     if (t > best) { best = t; *dx = -1.0f; *dy = -1.0f; }
    return (float)((double)(float)(best + c1 + c2 + c6) / 18.0);
  }
  float ddx = (float)(2 * c2 * c3 - c4 * c5) / (float)(-H);
  float ddy = (float)(2 * c1 * c4 - c3 * c5) / (float)(-H);
  const bool tx = ddx > 1.0f, tx_ = !tx && ddx < -1.0f, ty = ddy > 1.0f, ty_ = ddy < -1.0f;
  auto clamp1 = [](float v) { return v > 1.0f ? 1.0f : (v < -1.0f ? -1.0f : v); };
  // value of the fitted quadratic: int*float products are float, the sum is
  // accumulated left to right in float, the final division is double.
  auto quad = [&](float ax, float ay) {
    const float v = (float)c1 * ax * ax + (float)c2 * ay * ay + (float)c3 * ax + (float)c4 * ay +
                    (float)c5 * ax * ay + (float)c6;
    return (float)((double)v / 18.0);
  };
  if (tx || tx_ || ty || ty_) {
    float x1 = 0.0f, x2 = 0.0f, y1 = 0.0f, y2 = 0.0f;
    if (tx) { x1 = 1.0f; y1 = clamp1(-(float)(c4 + c5) / (float)(2 * c2)); }
    else if (tx_) { x1 = -1.0f; y1 = clamp1(-(float)(c4 - c5) / (float)(2 * c2)); }
    if (ty) { y2 = 1.0f; x2 = clamp1(-(float)(c3 + c5) / (float)(2 * c1)); }
    else if (ty_) { y2 = -1.0f; x2 = clamp1(-(float)(c3 - c5) / (float)(2 * c1)); }
    const float m1 = quad(x1, y1), m2 = quad(x2, y2);
    if (m1 > m2) { *dx = x1; *dy = x1; return m1; }
    *dx = x2; *dy = x2; return m2;
  }
  *dx = ddx; *dy = ddy;
  return quad(ddx, ddy);
}

// reference brisk-scale-space.cc:1101-1228: parabola through the scores at
// three scales.  kind 0 = Refine1D (0.75, 1, 1.5), 1 = Refine1D_1 (2/3, 1, 4/3),
// 2 = Refine1D_2 (0.7, 1, 1.5).
float Refine1D(int kind, float s_05, float s0, float s05, float* max) {
  const int i_05 = (int)(1024.0 * s_05 + 0.5), i0 = (int)(1024.0 * s0 + 0.5), i05 = (int)(1024.0 * s05 + 0.5);
  static const int kA[3][3] = {{16, -24, 8}, {9, -18, 9}, {2, -4, 2}};
  static const int kB[3][3] = {{-40, 54, -14}, {-21, 36, -15}, {-5, 8, -3}};
  static const int kC[3][3] = {{24, -27, 6}, {12, -16, 6}, {3, -3, 1}};
  static const double kLo[3] = {0.75, 0.6666666666666666666666666667, 0.7};
  static const double kHi[3] = {1.5, 1.3333333333333333333333333333, 1.5};
  static const double kDiv[3] = {3072.0, 2048.0, 1024.0};
  const int a = kA[kind][0] * i_05 + kA[kind][1] * i0 + kA[kind][2] * i05;
  if (a >= 0) {
    if (s0 >= s_05 && s0 >= s05) { *max = s0; return 1.0f; }
    if (s_05 >= s0 && s_05 >= s05) { *max = s_05; return (float)kLo[kind]; }
    if (s05 >= s0 && s05 >= s_05) { *max = s05; return (float)kHi[kind]; }
  }
  const int b = kB[kind][0] * i_05 + kB[kind][1] * i0 + kB[kind][2] * i05;
  float r = -(float)b / (float)(2 * a);
  if ((double)r < kLo[kind]) r = (float)kLo[kind];
  else if ((double)r > kHi[kind]) r = (float)kHi[kind];
  const int c = kC[kind][0] * i_05 + kC[kind][1] * i0 + kC[kind][2] * i05;
  float m = (float)c + (float)a * r * r + (float)b * r;
  if (kind == 2) m = m / 1024;  // `max /= 1024` (int literal -> float division)
  else m = (float)((double)m / kDiv[kind]);
  *max = m;
  return r;
}

struct ScaleSpace {
  std::vector<Layer> L;
  int threshold = 0;
  bool suppress = true;

  // reference brisk-scale-space.cc:64-90 (ConstructPyramid) and
  // brisk-layer.cc:53-95 (layer scale/offset).
  void Construct(const uint8_t* image, int w, int h, int octaves, int thresh) {
    threshold = thresh;
    const int n = octaves == 0 ? 1 : 2 * octaves;
    L.resize(n);
    L[0].w = w; L[0].h = h; L[0].img.assign(image, image + (size_t)w * h);
    L[0].scale = 1.0f; L[0].offset = 0.0f;
    for (int i = 1; i < n; ++i) {
      const Layer& src = (i == 1) ? L[0] : L[i - 2];
      Layer& d = L[i];
      if (i == 1) {
        d.w = 2 * (src.w / 3); d.h = 2 * (src.h / 3);
        d.img.resize((size_t)d.w * d.h);
        Twothirdsample8(src.img.data(), src.w, src.h, d.img.data());
        d.scale = (float)(src.scale * 1.5);
      } else {
        d.w = src.w / 2; d.h = src.h / 2;
        d.img.resize((size_t)d.w * d.h);
        Halfsample8(src.img.data(), src.w, src.h, d.img.data());
        d.scale = src.scale * 2;
      }
      d.offset = (float)(0.5 * d.scale - 0.5);
    }
    for (auto& l : L) l.Init(thresh);
  }

  // reference brisk-scale-space.cc:430-531 (IsMax2D).
  bool IsMax2D(int layer, int x, int y) {
    Layer& l = L[layer];
    const int W = l.w;
    const int center = l.scores[(size_t)y * W + x];
    static const int kOrder[8][2] = {{-1, 0}, {1, 0}, {0, -1}, {0, 1}, {-1, 1}, {1, 1}, {1, -1}, {-1, -1}};
    int s[8];
    for (int i = 0; i < 8; ++i) {
      s[i] = l.Score(x + kOrder[i][0], y + kOrder[i][1], center);
      if (center < s[i]) return false;
    }
    // ties: order (-1,-1) (0,-1) (1,-1) (-1,0) (1,0) (-1,1) (0,1) (1,1)
    static const int kTie[8] = {7, 2, 6, 0, 1, 4, 3, 5};
    const int smoothed = 4 * center + 2 * (s[0] + s[1] + s[2] + s[3]) + s[7] + s[6] + s[4] + s[5];
    for (int k = 0; k < 8; ++k) {
      const int i = kTie[k];
      if (s[i] != center) continue;
      const int cx = x + kOrder[i][0], cy = y + kOrder[i][1];
      const uint8_t* p = &l.scores[(size_t)(cy - 1) * W + cx - 1];
      const int other = p[0] + 2 * p[1] + p[2] + 2 * p[W] + 4 * p[W + 1] + 2 * p[W + 2] + p[2 * W] + 2 * p[2 * W + 1] +
                        p[2 * W + 2];
      if (other > smoothed) return false;
    }
    return true;
  }

  // Shared body of reference brisk-scale-space.cc:757-915 (GetScoreMaxAbove) and
  // :917-1099 (GetScoreMaxBelow): scan the projected patch [x_1,x1]x[y_1,y1] of
  // the neighbouring layer; reject above thr+5 (not on the bottom row); keep the
  // arg-max (the `below` scan has an extra tie rule on interior pixels).
  bool ScanPatch(Layer& nb, bool below, float x_1, float x1, float y_1, float y1, int threshold, float* max_out,
                 int* mx, int* my) {
    int max_x = (int)(x_1 + 1), max_y = (int)(y_1 + 1);
    float tmp;
    float max = (float)nb.ScoreF(x_1, y_1, 1);
    if (max > threshold) return false;
    for (int x = (int)(x_1 + 1); x <= (int)x1; ++x) {
      tmp = (float)nb.ScoreF((float)x, y_1, 1);
      if (tmp > threshold) return false;
      if (tmp > max) { max = tmp; max_x = x; }
    }
    tmp = (float)nb.ScoreF(x1, y_1, 1);
    if (tmp > threshold) return false;
    if (tmp > max) { max = tmp; max_x = (int)x1; }
    for (int y = (int)(y_1 + 1); y <= (int)y1; ++y) {
      tmp = (float)nb.ScoreF(x_1, (float)y, 1);
      if (tmp > threshold) return false;
      if (tmp > max) { max = tmp; max_x = (int)(x_1 + 1); max_y = y; }
      for (int x = (int)(x_1 + 1); x <= (int)x1; ++x) {
        tmp = (float)nb.Score(x, y, 1);
        if (tmp > threshold) return false;
        if (below && tmp == max) {
          const int t1 = 2 * (nb.Score(x - 1, y, 1) + nb.Score(x + 1, y, 1) + nb.Score(x, y + 1, 1) + nb.Score(x, y - 1, 1)) +
                         (nb.Score(x + 1, y + 1, 1) + nb.Score(x - 1, y + 1, 1) + nb.Score(x + 1, y - 1, 1) + nb.Score(x - 1, y - 1, 1));
          const int t2 = 2 * (nb.Score(max_x - 1, max_y, 1) + nb.Score(max_x + 1, max_y, 1) + nb.Score(max_x, max_y + 1, 1) + nb.Score(max_x, max_y - 1, 1)) +
                         (nb.Score(max_x + 1, max_y + 1, 1) + nb.Score(max_x - 1, max_y + 1, 1) + nb.Score(max_x + 1, max_y - 1, 1) + nb.Score(max_x - 1, max_y - 1, 1));
          if (t1 > t2) { max_x = x; max_y = y; }
        }
        if (tmp > max) { max = tmp; max_x = x; max_y = y; }
      }
      tmp = (float)nb.ScoreF(x1, (float)y, 1);
      if (tmp > threshold) return false;
      if (tmp > max) { max = tmp; max_x = (int)x1; max_y = y; }
    }
    tmp = (float)nb.ScoreF(x_1, y1, 1);
    if (tmp > max) { max = tmp; max_x = (int)(x_1 + 1); max_y = (int)y1; }
    for (int x = (int)(x_1 + 1); x <= (int)x1; ++x) {
      tmp = (float)nb.ScoreF((float)x, y1, 1);
      if (tmp > max) { max = tmp; max_x = x; max_y = (int)y1; }
    }
    tmp = (float)nb.ScoreF(x1, y1, 1);
    if (tmp > max) { max = tmp; max_x = (int)x1; max_y = (int)y1; }
    *max_out = max; *mx = max_x; *my = max_y;
    return true;
  }

  float Patch3x3(Layer& l, int x, int y, float* dx, float* dy, int* center = nullptr) {
    const int s00 = l.Score(x - 1, y - 1, 1), s10 = l.Score(x, y - 1, 1), s20 = l.Score(x + 1, y - 1, 1);
    const int s21 = l.Score(x + 1, y, 1), s11 = l.Score(x, y, 1), s01 = l.Score(x - 1, y, 1);
    const int s02 = l.Score(x - 1, y + 1, 1), s12 = l.Score(x, y + 1, 1), s22 = l.Score(x + 1, y + 1, 1);
    if (center) *center = s11;
    return Subpixel2D(s00, s01, s02, s10, s11, s12, s20, s21, s22, dx, dy);
  }

  // Same patch, but read through the float (bilinear) accessor: in
  // GetKeypoints the corner coordinates are `const float&`, so
  // `l.GetAgastScore(point_x - 1, point_y - 1, 1)` binds to the float overload
  // (brisk-layer.cc:147) and every read also touches (and caches) the pixels
  // at +1 in x and y.  Values are identical, the cache footprint is 4x4.
  float Patch3x3F(Layer& l, float x, float y, float* dx, float* dy) {
    const int s00 = l.ScoreF(x - 1, y - 1, 1), s10 = l.ScoreF(x, y - 1, 1), s20 = l.ScoreF(x + 1, y - 1, 1);
    const int s21 = l.ScoreF(x + 1, y, 1), s11 = l.ScoreF(x, y, 1), s01 = l.ScoreF(x - 1, y, 1);
    const int s02 = l.ScoreF(x - 1, y + 1, 1), s12 = l.ScoreF(x, y + 1, 1), s22 = l.ScoreF(x + 1, y + 1, 1);
    return Subpixel2D(s00, s01, s02, s10, s11, s12, s20, s21, s22, dx, dy);
  }

  static bool Saturate(float* dx, float* dy) {
    bool inside = true;
    if (*dx > 1.0f) { *dx = 1.0f; inside = false; }
    if (*dx < -1.0f) { *dx = -1.0f; inside = false; }
    if (*dy > 1.0f) { *dy = 1.0f; inside = false; }
    if (*dy < -1.0f) { *dy = -1.0f; inside = false; }
    return inside;
  }

  // reference brisk-scale-space.cc:757-915.
  float ScoreMaxAbove(int layer, int x, int y, int thr, bool* ismax, float* dx, float* dy) {
    *ismax = false;
    Layer& nb = L[layer + 1];
    float x_1, x1, y_1, y1;
    if (layer % 2 == 0) {
      x_1 = (float)((double)(float)(4 * x - 1 - 2) / 6.0); x1 = (float)((double)(float)(4 * x - 1 + 2) / 6.0);
      y_1 = (float)((double)(float)(4 * y - 1 - 2) / 6.0); y1 = (float)((double)(float)(4 * y - 1 + 2) / 6.0);
    } else {
      x_1 = (float)(6 * x - 1 - 3) / 8.0f; x1 = (float)(6 * x - 1 + 3) / 8.0f;
      y_1 = (float)(6 * y - 1 - 3) / 8.0f; y1 = (float)(6 * y - 1 + 3) / 8.0f;
    }
    float max; int mx, my;
    if (!ScanPatch(nb, false, x_1, x1, y_1, y1, thr + 5, &max, &mx, &my)) return 0.0f;
    float dx1, dy1;
    const float refined = Patch3x3(nb, mx, my, &dx1, &dy1);
    const float rx = (float)mx + dx1, ry = (float)my + dy1;
    if (layer % 2 == 0) {
      *dx = (rx * 6.0f + 1.0f) / 4.0f - (float)x;
      *dy = (ry * 6.0f + 1.0f) / 4.0f - (float)y;
    } else {
      *dx = (float)(((double)rx * 8.0 + 1.0) / 6.0 - (double)(float)x);
      *dy = (float)(((double)ry * 8.0 + 1.0) / 6.0 - (double)(float)y);
    }
    const bool inside = Saturate(dx, dy);
    *ismax = true;
    return inside ? std::max(refined, max) : max;
  }

  // reference brisk-scale-space.cc:917-1099.
  float ScoreMaxBelow(int layer, int x, int y, int thr, bool* ismax, float* dx, float* dy) {
    *ismax = false;
    Layer& nb = L[layer - 1];
    float x_1, x1, y_1, y1;
    if (layer % 2 == 0) {
      x_1 = (float)((double)(float)(8 * x + 1 - 4) / 6.0); x1 = (float)((double)(float)(8 * x + 1 + 4) / 6.0);
      y_1 = (float)((double)(float)(8 * y + 1 - 4) / 6.0); y1 = (float)((double)(float)(8 * y + 1 + 4) / 6.0);
    } else {
      x_1 = (float)((double)(float)(6 * x + 1 - 3) / 4.0); x1 = (float)((double)(float)(6 * x + 1 + 3) / 4.0);
      y_1 = (float)((double)(float)(6 * y + 1 - 3) / 4.0); y1 = (float)((double)(float)(6 * y + 1 + 3) / 4.0);
    }
    float max; int mx, my;
    if (!ScanPatch(nb, true, x_1, x1, y_1, y1, thr + 5, &max, &mx, &my)) return 0.0f;
    float dx1, dy1;
    const float refined = Patch3x3(nb, mx, my, &dx1, &dy1);
    const float rx = (float)mx + dx1, ry = (float)my + dy1;
    if (layer % 2 == 0) {
      *dx = (float)(((double)rx * 6.0 + 1.0) / 8.0 - (double)(float)x);
      *dy = (float)(((double)ry * 6.0 + 1.0) / 8.0 - (double)(float)y);
    } else {
      *dx = (float)(((double)rx * 4.0 - 1.0) / 6.0 - (double)(float)x);
      *dy = (float)(((double)ry * 4.0 - 1.0) / 6.0 - (double)(float)y);
    }
    const bool inside = Saturate(dx, dy);
    *ismax = true;
    return inside ? std::max(refined, max) : max;
  }

  // reference brisk-scale-space.cc:534-754 (Refine3D).
  float Refine3D(int layer, int x_layer, int y_layer, float* x, float* y, float* scale, bool* ismax) {
    *ismax = true;
    Layer& l = L[layer];
    const int center = l.Score(x_layer, y_layer, 1);
    float dxa = 0, dya = 0;
    const float max_above = ScoreMaxAbove(layer, x_layer, y_layer, center, ismax, &dxa, &dya);
    if (!*ismax) return 0.0f;
    float max = 0;
    bool refine_scale = true;
    float dxb = 0, dyb = 0, max_below;
    const bool octave = layer % 2 == 0;
    if (layer == 0) {
      // guess the virtual intra-octave below octave 0 with the 5-8 mask (:558-592)
      int best = 0;
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) best = std::max(best, l.Score58(x_layer + dx, y_layer + dy, 1));
      const int s00 = l.Score58(x_layer - 1, y_layer - 1, 1), s10 = l.Score58(x_layer, y_layer - 1, 1);
      const int s20 = l.Score58(x_layer + 1, y_layer - 1, 1), s21 = l.Score58(x_layer + 1, y_layer, 1);
      const int s11 = l.Score58(x_layer, y_layer, 1), s01 = l.Score58(x_layer - 1, y_layer, 1);
      const int s02 = l.Score58(x_layer - 1, y_layer + 1, 1), s12 = l.Score58(x_layer, y_layer + 1, 1);
      const int s22 = l.Score58(x_layer + 1, y_layer + 1, 1);
      Subpixel2D(s00, s01, s02, s10, s11, s12, s20, s21, s22, &dxb, &dyb);
      max_below = (float)best;
    } else {
      max_below = ScoreMaxBelow(layer, x_layer, y_layer, center, ismax, &dxb, &dyb);
      if (!*ismax) return 0.0f;
    }
    float dxl, dyl; int s11;
    const float max_layer = Patch3x3(l, x_layer, y_layer, &dxl, &dyl, &s11);
    // scale-axis acceptance (:611-629, :701-713); kMaxThreshold_=1, kMinDrop_=15
    if (layer == 0) {
      if (s11 - 1 <= (int)max_above) refine_scale = false;
    } else if ((float)(s11 - 1) < max_above || (float)(s11 - 1) < max_below) {
      if ((float)(s11 - 15) > max_above || (float)(s11 - 15) > max_below) refine_scale = false;
      else { *ismax = false; return 0.0f; }
    }
    if (refine_scale) {
      const int kind = octave ? (layer == 0 ? 2 : 0) : 1;
      *scale = Refine1D(kind, max_below, std::max((float)center, max_layer), max_above, &max);
    } else {
      *scale = 1.0f;
      max = max_layer;
    }
    float r0, r1, ox, oy;  // interpolation weights and the other layer's offsets
    if (octave) {
      if (*scale > 1.0f) { r0 = (float)((1.5 - (double)*scale) / .5); ox = dxa; oy = dya; }
      else if (layer == 0) { r0 = (float)(((double)*scale - 0.5) / 0.5); ox = dxb; oy = dyb; }
      else { r0 = (float)(((double)*scale - 0.75) / 0.25); ox = dxb; oy = dyb; }
    } else {
      if (*scale > 1.0f) { r0 = (float)(4.0 - (double)*scale * 3.0); ox = dxa; oy = dya; }
      else { r0 = (float)((double)*scale * 3.0 - 2.0); ox = dxb; oy = dyb; }
    }
    r1 = (float)(1.0 - (double)r0);
    const float px = r0 * dxl + r1 * ox + (float)x_layer;
    const float py = r0 * dyl + r1 * oy + (float)y_layer;
    if (layer == 0 && !(*scale > 1.0f)) { *x = px; *y = py; }
    else { *x = px * l.scale + l.offset; *y = py * l.scale + l.offset; }
    *scale *= l.scale;
    return max;
  }

  // reference brisk-scale-space.cc:92-287 with a non-empty key-point vector ("provided key points",
  // :104-124; reached through BriskFeatureDetector::ComputeScale, brisk-feature-detector.cc:87-92, which
  // builds the pyramid with lower threshold 0).  Per layer the points are mapped into layer coordinates and
  // kept when inside the 3-pixel border; their four bilinear neighbours are scored with threshold 0; the
  // layer's detector only runs when no point was kept (brisk-layer.cc:103-105); every point then gets the
  // score cornerScore(b = thrmap) written at the flat offset int(x + y * cols) -- evaluated in float, so a
  // fractional y lands on some other column (brisk-layer.cc:110-116).  No 2-D non-maximum test afterwards.
  void GetKeypointsProvided(const orc_keypoint* in, int n_in, std::vector<orc_keypoint>* out) {
    const int n = (int)L.size();
    struct P { float x, y; int class_id; };
    std::vector<std::vector<P>> pts(n);
    for (int i = 0; i < n; ++i) {
      Layer& l = L[i];
      for (int k = 0; k < n_in; ++k) {
        const float x = in[k].x / l.scale - l.offset, y = in[k].y / l.scale - l.offset;
        if (x < 3 || y < 3 || x > l.w - 3 || y > l.h - 3) continue;
        l.ScoreF(x, y, 0);
        pts[i].push_back(P{x, y, in[k].class_id});
      }
      if (pts[i].empty()) {
        std::vector<Corner> c;
        DetectCorners(l.img.data(), l.thr.data(), l.w, l.h, threshold, 0, 230, &c);
        for (const Corner& q : c) pts[i].push_back(P{(float)q.x, (float)q.y, -1});
      }
      for (const P& p : pts[i]) {
        const int offs = (int)(p.x + p.y * (float)l.w);
        const int qx = offs % l.w, qy = offs / l.w;
        // cornerScore(b = thrmap) = max(thrmap, F) = thrmap wherever the ring is inside the image; border
        // pixels (where the reference reads its ring across row ends) are never read back (Score())
        if (qx >= 3 && qy >= 3 && qx < l.w - 3 && qy < l.h - 3) l.scores[(size_t)offs] = l.thr[(size_t)offs];
      }
    }
    out->clear();
    auto emit = [&](float x, float y, float size, float response, int octave, int class_id) {
      out->push_back(orc_keypoint{x, y, size, -1.0f, response, octave, class_id});
    };
    if (n == 1) {  // :172-209, and :131-170 when scale non-maxima are kept (same arithmetic for one layer)
      for (const P& p : pts[0]) {
        float dx, dy;
        const float max = Patch3x3F(L[0], p.x, p.y, &dx, &dy);
        emit(p.x + dx, p.y + dy, 12.0f, max, 0, p.class_id);
      }
      return;
    }
    for (int i = 0; i < n; ++i) {
      Layer& l = L[i];
      for (const P& p : pts[i]) {
        if (i == n - 1) {  // :215-256
          bool ismax; float dx, dy;
          ScoreMaxBelow(i, (int)p.x, (int)p.y, l.ScoreF(p.x, p.y, 1), &ismax, &dx, &dy);
          if (!ismax) continue;
          float ddx, ddy;
          const float max = Patch3x3F(l, p.x, p.y, &ddx, &ddy);
          emit((p.x + ddx) * l.scale + l.offset, (p.y + ddy) * l.scale + l.offset, 12.0f * l.scale, max, i, p.class_id);
        } else {  // :257-285
          bool ismax; float x, y, scale;
          const float score = Refine3D(i, (int)p.x, (int)p.y, &x, &y, &scale, &ismax);
          if (!ismax) continue;
          emit(x, y, 12.0f * scale, score, i, p.class_id);
        }
      }
    }
  }

  // reference brisk-scale-space.cc:92-287 (GetKeypoints, detection mode).
  void GetKeypoints(std::vector<orc_keypoint>* out) {
    const int n = (int)L.size();
    std::vector<std::vector<Corner>> pts(n);
    for (int i = 0; i < n; ++i) {
      DetectCorners(L[i].img.data(), L[i].thr.data(), L[i].w, L[i].h, threshold, 10, 230, &pts[i]);
      for (const Corner& c : pts[i]) L[i].scores[(size_t)c.y * L[i].w + c.x] = (uint8_t)c.score;
    }
    out->clear();
    auto emit = [&](float x, float y, float size, float response, int octave) {
      out->push_back(orc_keypoint{x, y, size, -1.0f, response, octave, -1});
    };
    if (!suppress) {
      // :131-170 -- note the reference indexes agastPoints.at(0)[n] for every
      // layer; reproduced.
      for (int i = 0; i < n; ++i) {
        const int num = (int)pts[i].size();
        for (int k = 0; k < num; ++k) {
          if (k >= (int)pts[0].size()) break;  // the reference would read out of bounds here
          const Corner& c = pts[0][k];
          if (!IsMax2D(i, c.x, c.y)) continue;
          float dx, dy;
          const float max = Patch3x3F(L[i], (float)c.x, (float)c.y, &dx, &dy);
          emit((float)c.x + dx, (float)c.y + dy, 12.0f * L[i].scale, max, 0);
        }
      }
      return;
    }
    if (n == 1) {  // :172-209
      for (const Corner& c : pts[0]) {
        if (!IsMax2D(0, c.x, c.y)) continue;
        float dx, dy;
        const float max = Patch3x3F(L[0], (float)c.x, (float)c.y, &dx, &dy);
        emit((float)c.x + dx, (float)c.y + dy, 12.0f, max, 0);
      }
      return;
    }
    for (int i = 0; i < n; ++i) {
      Layer& l = L[i];
      for (const Corner& c : pts[i]) {
        if (!IsMax2D(i, c.x, c.y)) continue;
        if (i == n - 1) {  // :215-256
          bool ismax; float dx, dy;
          ScoreMaxBelow(i, c.x, c.y, l.ScoreF((float)c.x, (float)c.y, 1), &ismax, &dx, &dy);
          if (!ismax) continue;
          float ddx, ddy;
          const float max = Patch3x3F(l, (float)c.x, (float)c.y, &ddx, &ddy);
          emit(((float)c.x + ddx) * l.scale + l.offset, ((float)c.y + ddy) * l.scale + l.offset, 12.0f * l.scale, max, i);
        } else {  // :257-285
          bool ismax; float x, y, scale;
          const float score = Refine3D(i, c.x, c.y, &x, &y, &scale, &ismax);
          if (!ismax) continue;
          emit(x, y, 12.0f * scale, score, i);
        }
      }
    }
  }
};

// ---------------------------------------------------------------------------
// Integral image + descriptor.
// ---------------------------------------------------------------------------

// reference integral-image.h:56-161: S(y+1,x+1) = sum_{v<=y,u<=x} I(v,u), zero
// first row and column, int32.
void Integral8(const uint8_t* img, int w, int h, int32_t* out) {
  const int sw = w + 1;
  memset(out, 0, sizeof(int32_t) * sw);
  for (int y = 0; y < h; ++y) {
    int32_t s = 0;
    out[(size_t)(y + 1) * sw] = 0;
    for (int x = 0; x < w; ++x) {
      s += img[(size_t)y * w + x];
      out[(size_t)(y + 1) * sw + x + 1] = out[(size_t)y * sw + x + 1] + s;
    }
  }
}

struct PatternPoint { float x, y, sigma; };
struct LongPair { unsigned i, j; int wdx, wdy; };
struct ShortPair { unsigned i, j; };

struct Pattern {
  unsigned points = 0;
  std::vector<PatternPoint> pts;  // [64][1024][points]
  float scale_list[64];
  unsigned size_list[64];
  std::vector<ShortPair> shorts;
  std::vector<LongPair> longs;
  int strings = 0;

  static const unsigned kScales = 64, kRot = 1024;

  static float LbScaleStep() {
    // reference brisk-descriptor-extractor.cc:85-86 / :200-201
    static const float lb_scale = (float)(log((double)30.0f) / log(2.0));
    static const float lb_scale_step = lb_scale / (kScales);
    return lb_scale_step;
  }

  // reference brisk-descriptor-extractor.cc:180-291 (InitFromStream) on the
  // default pattern of pattern-provider.cc (data in brisk_pattern_data.inc).
  void InitV2(float pattern_scale) {
    points = kBrisk2NumPoints;
    std::vector<float> ux(points), uy(points), sg(points);
    for (unsigned i = 0; i < points; ++i) {
      float f[3];
      memcpy(f, &kBrisk2PointBits[3 * i], sizeof(f));
      ux[i] = f[0] * pattern_scale; uy[i] = f[1] * pattern_scale; sg[i] = f[2] * pattern_scale;
    }
    pts.resize((size_t)kScales * kRot * points);
    const float sigma_scale = 1.3f;
    PatternPoint* it = pts.data();
    for (unsigned s = 0; s < kScales; ++s) {
      scale_list[s] = (float)pow(2.0, (double)(s * LbScaleStep()));
      size_list[s] = 0;
      for (unsigned rot = 0; rot < kRot; ++rot) {
        const double theta = (double)rot * 2 * M_PI / (double)kRot;
        for (unsigned i = 0; i < points; ++i, ++it) {
          it->x = (float)(scale_list[s] * (ux[i] * cos(theta) - uy[i] * sin(theta)));
          it->y = (float)(scale_list[s] * (ux[i] * sin(theta) + uy[i] * cos(theta)));
          it->sigma = sigma_scale * scale_list[s] * sg[i];
          const unsigned size = (unsigned)(ceil(sqrt((double)(it->x * it->x + it->y * it->y)) + it->sigma) + 1);
          if (size_list[s] < size) size_list[s] = size;
        }
      }
    }
    shorts.resize(kBrisk2NumShortPairs);
    for (int p = 0; p < kBrisk2NumShortPairs; ++p) { shorts[p].i = kBrisk2ShortPairs[2 * p]; shorts[p].j = kBrisk2ShortPairs[2 * p + 1]; }
    longs.resize(kBrisk2NumLongPairs);
    for (int p = 0; p < kBrisk2NumLongPairs; ++p) {
      const unsigned i = kBrisk2LongPairs[2 * p], j = kBrisk2LongPairs[2 * p + 1];
      const float dx = ux[j] - ux[i], dy = uy[j] - uy[i];
      const float nsq = dx * dx + dy * dy;
      longs[p] = LongPair{i, j, (int)((dx / nsq) * 2048.0 + 0.5), (int)((dy / nsq) * 2048.0 + 0.5)};
    }
    strings = (int)ceil((float)shorts.size() / 128.0) * 4 * 4;
  }

  // reference brisk-descriptor-extractor.cc:65-178 (generateKernel) with the
  // BRISK 1.0 rings of :316-339.
  void InitV1(float pattern_scale) {
    const double f = 0.85 * pattern_scale;
    const float radius[5] = {(float)(f * 0), (float)(f * 2.9), (float)(f * 4.9), (float)(f * 7.4), (float)(f * 10.8)};
    const int number[5] = {1, 10, 14, 15, 20};
    const float dMax = 5.85f, dMin = 8.2f;
    points = 0;
    for (int r = 0; r < 5; ++r) points += number[r];
    pts.resize((size_t)kScales * kRot * points);
    const float sigma_scale = 1.3f;
    PatternPoint* it = pts.data();
    for (unsigned s = 0; s < kScales; ++s) {
      scale_list[s] = (float)pow((double)2.0, (double)(s * LbScaleStep()));
      size_list[s] = 0;
      for (unsigned rot = 0; rot < kRot; ++rot) {
        const double theta = (double)rot * 2 * M_PI / (double)kRot;
        for (int ring = 0; ring < 5; ++ring)
          for (int num = 0; num < number[ring]; ++num, ++it) {
            const double alpha = ((double)num) * 2 * M_PI / (double)number[ring];
            it->x = (float)(scale_list[s] * radius[ring] * cos(alpha + theta));
            it->y = (float)(scale_list[s] * radius[ring] * sin(alpha + theta));
            if (ring == 0) it->sigma = (float)(sigma_scale * scale_list[s] * 0.5);
            else it->sigma = (float)(sigma_scale * scale_list[s] * ((double)radius[ring]) * sin(M_PI / number[ring]));
            const unsigned size = (unsigned)(ceil((scale_list[s] * radius[ring]) + it->sigma) + 1);
            if (size_list[s] < size) size_list[s] = size;
          }
      }
    }
    const float dmin_sq = dMin * dMin, dmax_sq = dMax * dMax;
    for (unsigned i = 1; i < points; ++i)
      for (unsigned j = 0; j < i; ++j) {
        const float dx = pts[j].x - pts[i].x, dy = pts[j].y - pts[i].y;
        const float nsq = dx * dx + dy * dy;
        if (nsq > dmin_sq) longs.push_back(LongPair{i, j, (int)((dx / nsq) * 2048.0 + 0.5), (int)((dy / nsq) * 2048.0 + 0.5)});
        if (nsq < dmax_sq) shorts.push_back(ShortPair{i, j});
      }
    strings = (int)ceil((float)shorts.size() / 128.0) * 4 * 4;
  }
};

std::mutex g_mu;
std::map<std::pair<int, float>, std::shared_ptr<Pattern>> g_patterns;

std::shared_ptr<Pattern> GetPattern(int version, float pattern_scale) {
  std::lock_guard<std::mutex> lock(g_mu);
  auto key = std::make_pair(version, pattern_scale);
  auto it = g_patterns.find(key);
  if (it != g_patterns.end()) return it->second;
  std::shared_ptr<Pattern> p(new Pattern);
  if (version == 1) p->InitV1(pattern_scale); else p->InitV2(pattern_scale);
  g_patterns[key] = p;
  return p;
}

// reference brisk-descriptor-extractor.cc:370-530 (SmoothedIntensity<uchar,int>),
// SURVEY.md App. A.6.  Box-filtered intensity (scaled by ~1024) around a pattern
// point: bilinear for sigma<0.5, else area-weighted with the integral image for
// large boxes -- including the reference's skewed bottom-corner reads (:447-456).
int SmoothedIntensity(const uint8_t* img, int w, const int32_t* integral, float kx, float ky, const PatternPoint& pp) {
  const float xf = pp.x + kx, yf = pp.y + ky;
  const int x = (int)xf, y = (int)yf;
  const float sigma_half = pp.sigma;
  const float area = (float)(4.0 * sigma_half * sigma_half);
  if (sigma_half < 0.5) {
    const int r_x = (int)((xf - x) * 1024), r_y = (int)((yf - y) * 1024);
    const int r_x_1 = 1024 - r_x, r_y_1 = 1024 - r_y;
    const uint8_t* p = img + x + (size_t)y * w;
    int v = r_x_1 * r_y_1 * (int)p[0];
    v += r_x * r_y_1 * (int)p[1];
    v += r_x * r_y * (int)p[w + 1];
    v += r_x_1 * r_y * (int)p[w];
    return v / 1024;
  }
  const int scaling = (int)(4194304.0 / area);
  const int scaling2 = (int)((float)scaling * area / 1024.0);
  const int iw = w + 1;
  const float x_1 = xf - sigma_half, x1 = xf + sigma_half, y_1 = yf - sigma_half, y1 = yf + sigma_half;
  const int x_left = (int)(x_1 + 0.5), y_top = (int)(y_1 + 0.5), x_right = (int)(x1 + 0.5), y_bottom = (int)(y1 + 0.5);
  const float r_x_1 = (float)((float)x_left - x_1 + 0.5), r_y_1 = (float)((float)y_top - y_1 + 0.5);
  const float r_x1 = (float)(x1 - (float)x_right + 0.5), r_y1 = (float)(y1 - (float)y_bottom + 0.5);
  const int dx = x_right - x_left - 1, dy = y_bottom - y_top - 1;
  const int A = (int)((r_x_1 * r_y_1) * scaling), B = (int)((r_x1 * r_y_1) * scaling);
  const int C = (int)((r_x1 * r_y1) * scaling), D = (int)((r_x_1 * r_y1) * scaling);
  const int r_x_1_i = (int)(r_x_1 * scaling), r_y_1_i = (int)(r_y_1 * scaling);
  const int r_x1_i = (int)(r_x1 * scaling), r_y1_i = (int)(r_y1 * scaling);
  if (dx + dy > 2) {
    const uint8_t* p = img + x_left + (size_t)w * y_top;
    int v = A * (int)p[0];
    p += dx + 1;            v += B * (int)p[0];
    p += dy * w + 1;        v += C * (int)p[0];
    p -= dx + 1;            v += D * (int)p[0];
    const int32_t* q = integral + x_left + (size_t)iw * y_top + 1;
    const int t1 = *q;  q += dx;
    const int t2 = *q;  q += iw;
    const int t3 = *q;  q++;
    const int t4 = *q;  q += dy * iw;
    const int t5 = *q;  q--;
    const int t6 = *q;  q += iw;
    const int t7 = *q;  q -= dx;
    const int t8 = *q;  q -= iw;
    const int t9 = *q;  q--;
    const int t10 = *q; q -= dy * iw;
    const int t11 = *q; q++;
    const int t12 = *q;
    const int upper = (t3 - t2 + t1 - t12) * r_y_1_i;
    const int middle = (t6 - t3 + t12 - t9) * scaling;
    const int left = (t9 - t12 + t11 - t10) * r_x_1_i;
    const int right = (t5 - t4 + t3 - t6) * r_x1_i;
    const int bottom = (t7 - t6 + t9 - t8) * r_y1_i;
    return (v + upper + middle + left + right + bottom) / scaling2;
  }
  // explicit weighted sum over the window (:497-529), as an index walk: when
  // float rounding makes the window degenerate (dx or dy == -1 for 2*sigma ~ 1)
  // the reference's pointer arithmetic revisits / drifts over neighbouring
  // pixels, which only the literal walk reproduces.
  const uint8_t* p = img + x_left + (size_t)w * y_top;
  long i = 0;
  int v = A * (int)p[i];
  ++i;
  for (long e = i + dx; i < e; ++i) v += r_y_1_i * (int)p[i];
  v += B * (int)p[i];
  i += w - dx - 1;
  for (long ej = i + (long)dy * w; i < ej; i += w - dx - 1) {
    v += r_x_1_i * (int)p[i];
    ++i;
    for (long e = i + dx; i < e; ++i) v += (int)p[i] * scaling;
    v += r_x1_i * (int)p[i];
  }
  v += D * (int)p[i];
  ++i;
  for (long e = i + dx; i < e; ++i) v += r_y1_i * (int)p[i];
  v += C * (int)p[i];
  return v / scaling2;
}

// reference brisk-descriptor-extractor.cc:612-778 (doDescriptorComputation).
int Describe(const uint8_t* img, int w, int h, orc_keypoint* kps, int n, bool rot_inv, bool scale_inv,
             const Pattern& pat, uint8_t* desc) {
  static const float log2f_ = 0.693147180559945f;
  static const float lb_scalerange = (float)(log((double)30.0f) / (double)(log2f_));
  static const float basicSize06 = (float)(12.0f * 0.6);
  unsigned basicscale = 0;
  if (!scale_inv)
    basicscale = std::max((int)(Pattern::kScales / lb_scalerange * (log(1.45 * 12.0f / (basicSize06)) / (double)log2f_) + 0.5), 0);
  std::vector<orc_keypoint> valid;
  std::vector<int> scales;
  for (int k = 0; k < n; ++k) {
    unsigned scale;
    if (scale_inv) {
      scale = std::max((int)(Pattern::kScales / lb_scalerange * (log((double)(kps[k].size / (basicSize06))) / (double)log2f_) + 0.5), 0);
      if (scale >= Pattern::kScales) scale = Pattern::kScales - 1;
    } else {
      scale = basicscale;
    }
    const int border = pat.size_list[scale];
    const int bx = w - border, by = h - border;
    const bool outside = (kps[k].x < (float)border) || (kps[k].x >= (float)bx) || (kps[k].y < (float)border) ||
                         (kps[k].y >= (float)by);
    if (!outside) { valid.push_back(kps[k]); scales.push_back((int)scale); }
  }
  const int nv = (int)valid.size();
  memset(desc, 0, (size_t)nv * pat.strings);
  std::vector<int32_t> integral((size_t)(w + 1) * (h + 1));
  Integral8(img, w, h, integral.data());
  std::vector<int> values(pat.points);
  const unsigned P = pat.points;
  for (int k = 0; k < nv; ++k) {
    orc_keypoint& kp = valid[k];
    const int scale = scales[k];
    int theta;
    if (kp.angle == -1) {
      if (!rot_inv) {
        theta = 0;
      } else {
        const PatternPoint* pp = &pat.pts[((size_t)scale * Pattern::kRot + 0) * P];
        for (unsigned i = 0; i < P; ++i) values[i] = SmoothedIntensity(img, w, integral.data(), kp.x, kp.y, pp[i]);
        int d0 = 0, d1 = 0;
        for (const LongPair& lp : pat.longs) {
          const int delta = values[lp.i] - values[lp.j];
          d0 += delta * lp.wdx / 1024;
          d1 += delta * lp.wdy / 1024;
        }
        kp.angle = (float)(atan2((double)(float)d1, (double)(float)d0) / M_PI * 180.0);
        theta = (int)((Pattern::kRot * kp.angle) / (360.0) + 0.5);
        if (theta < 0) theta += Pattern::kRot;
        if (theta >= (int)Pattern::kRot) theta -= Pattern::kRot;
      }
    } else {
      if (!rot_inv) {
        theta = 0;
      } else {
        theta = (int)(Pattern::kRot * (kp.angle / (360.0)) + 0.5);
        if (theta < 0) theta += Pattern::kRot;
        if (theta >= (int)Pattern::kRot) theta -= Pattern::kRot;
      }
    }
    const PatternPoint* pp = &pat.pts[((size_t)scale * Pattern::kRot + theta) * P];
    for (unsigned i = 0; i < P; ++i) values[i] = SmoothedIntensity(img, w, integral.data(), kp.x, kp.y, pp[i]);
    // reference :538-564 (setDescriptorBits): bit p%32 of little-endian word p/32
    uint32_t* words = reinterpret_cast<uint32_t*>(desc + (size_t)k * pat.strings);
    for (size_t p = 0; p < pat.shorts.size(); ++p)
      if (values[pat.shorts[p].i] > values[pat.shorts[p].j]) words[p / 32] |= 1u << (p % 32);
  }
  for (int k = 0; k < nv; ++k) kps[k] = valid[k];
  return nv;
}

// ---------------------------------------------------------------------------
// Harris scale space.
// ---------------------------------------------------------------------------

// reference harris-scores.cc:53-279 (HarrisScoresSSE), SURVEY.md App. A.5.
void HarrisScores(const uint8_t* img, int w, int h, int32_t* out) {
  memset(out, 0, sizeof(int32_t) * w * h);
  if (w < 5 || h < 5) return;
  std::vector<int16_t> xx((size_t)w * h, 0), yy((size_t)w * h, 0), xy((size_t)w * h, 0);
  auto I = [&](int y, int x) { return (int)img[(size_t)y * w + x]; };
  for (int y = 1; y < h - 1; ++y)
    for (int x = 1; x < w - 1; ++x) {
      const int dx = (10 * (I(y, x - 1) - I(y, x + 1)) + 3 * (I(y - 1, x - 1) - I(y - 1, x + 1)) + 3 * (I(y + 1, x - 1) - I(y + 1, x + 1))) * 8;
      const int dy = (10 * (I(y - 1, x) - I(y + 1, x)) + 3 * (I(y - 1, x - 1) - I(y + 1, x - 1)) + 3 * (I(y - 1, x + 1) - I(y + 1, x + 1))) * 8;
      const int16_t sdx = (int16_t)dx, sdy = (int16_t)dy;
      xx[(size_t)y * w + x] = (int16_t)(((int)sdx * sdx) >> 16);
      yy[(size_t)y * w + x] = (int16_t)(((int)sdy * sdy) >> 16);
      xy[(size_t)y * w + x] = (int16_t)(((int)sdx * sdy) >> 16);
    }
  auto G = [&](const std::vector<int16_t>& p, int y, int x) {
    const int16_t* q = &p[(size_t)y * w + x];
    const int s = 4 * q[0] + 2 * (q[-w] + q[w] + q[-1] + q[1]) + q[-w - 1] + q[-w + 1] + q[w - 1] + q[w + 1];
    return s >> 4;
  };
  for (int y = 2; y < h - 2; ++y)
    for (int x = 2; x < w - 2; ++x) {
      const int a = G(xx, y, x), b = G(yy, y, x), c = G(xy, y, x);
      const int t = (a + b) >> 1;
      out[(size_t)y * w + x] = a * b - c * c - ((t * t) >> 2);
    }
}

struct ScoredPoint {
  int score; uint16_t x, y;
  bool operator<(const ScoredPoint& o) const { return score > o.score; }  // score-calculator.h:82-84
};

// reference harris-score-calculator.cc:57-106 (Get2dMaxima).
void HarrisMaxima(const int32_t* sc, int w, int h, int abs_thr, std::vector<ScoredPoint>* out) {
  out->clear();
  for (int y = 2; y < h - 2; ++y)
    for (int x = 2; x < w - 2; ++x) {
      const int32_t* p = sc + (size_t)y * w + x;
      const int c = *p;
      if (c < abs_thr) continue;
      if (p[1] > c || p[-1] > c || p[w] > c || p[-w] > c || p[w + 1] > c || p[w - 1] > c || p[-w + 1] > c || p[-w - 1] > c) continue;
      out->push_back(ScoredPoint{c, (uint16_t)x, (uint16_t)y});
    }
}

struct HarrisLayer {
  int w = 0, h = 0, number = 0;
  bool octave = true;
  Bytes img;
  std::vector<int32_t> scores;
  double offset_above, scale_above, offset_below, scale_below, scale, offset;

  // reference harris-score-calculator.h:57-74 (bilinear double read, 0 outside)
  double ScoreD(double u, double v) const {
    const int ui = (int)u, vi = (int)v;
    if (ui + 1 >= w || vi + 1 >= h || ui < 0 || vi < 0) return 0.0;
    const double ru = u - (double)ui, rv = v - (double)vi;
    const double mu = 1.0 - ru, mv = 1.0 - rv;
    const int32_t* p = &scores[(size_t)vi * w + ui];
    return mv * (mu * p[0] + ru * p[1]) + rv * (mu * p[w] + ru * p[w + 1]);
  }
};

// reference scale-space-layer-inl.h:559-693 (Subpixel2D, double inputs); only
// delta_x / delta_y are consumed by the caller.
void HarrisSubpixel2D(double s00, double s01, double s02, double s10, double s11, double s12, double s20, double s21,
                      double s22, float* dx, float* dy) {
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
    int best = (int)(c3 + c4 + c5); *dx = 1.0f; *dy = 1.0f;
    int t = (int)(-c3 + c4 - c5); if (t > best) { best = t; *dx = -1.0f; *dy = 1.0f; }
    t = (int)(c3 - c4 - c5);      if (t > best) { best = t; *dx = 1.0f; *dy = -1.0f; }
    t = (int)(-c3 - c4 + c5);     if (t > best) { best = t; *dx = -1.0f; *dy = -1.0f; }
    return;
  }
  const float ddx = (float)(2 * c2 * c3 - c4 * c5) / (float)(-H);
  const float ddy = (float)(2 * c1 * c4 - c3 * c5) / (float)(-H);
  const bool tx = ddx > 1.0f, tx_ = !tx && ddx < -1.0f, ty = ddy > 1.0f, ty_ = ddy < -1.0f;
  auto clamp1 = [](float v) { return v > 1.0f ? 1.0f : (v < -1.0f ? -1.0f : v); };
  if (tx || tx_ || ty || ty_) {
    float x1 = 0.0f, x2 = 0.0f, y1 = 0.0f, y2 = 0.0f;
    if (tx) { x1 = 1.0f; y1 = clamp1(-(float)(c4 + c5) / (float)(2 * c2)); }
    else if (tx_) { x1 = -1.0f; y1 = clamp1(-(float)(c4 - c5) / (float)(2 * c2)); }
    if (ty) { y2 = 1.0f; x2 = clamp1(-(float)(c3 + c5) / (float)(2 * c1)); }
    else if (ty_) { y2 = -1.0f; x2 = clamp1(-(float)(c3 - c5) / (float)(2 * c1)); }
    auto quad = [&](float ax, float ay) {
      return (float)((c1 * ax * ax + c2 * ay * ay + c3 * ax + c4 * ay + c5 * ax * ay + c6) / 18.0);
    };
    const float m1 = quad(x1, y1), m2 = quad(x2, y2);
    if (m1 > m2) { *dx = x1; *dy = x1; } else { *dx = x2; *dy = x2; }
    return;
  }
  *dx = ddx; *dy = ddy;
}

// reference uniformity-enforcement-inl.h:44-194.
void EnforceUniformity(double radius, int rows, int cols, size_t max_kpt, std::vector<ScoredPoint>* points) {
  float lut[31][31];  // scale-space-layer-inl.h:89-97
  for (int x = 0; x < 31; ++x)
    for (int y = 0; y < 31; ++y)
      lut[y][x] = (float)std::max(1 - (double)((15 - x) * (15 - x) + (15 - y) * (15 - y)) / (double)(15 * 15), 0.0);
  std::sort(points->begin(), points->end());
  const float max_score = (float)points->front().score;
  const float scaling = (float)(15.0 / (float)radius);
  const int orows = (int)(rows * ceil(scaling) + 32), ocols = (int)(cols * ceil(scaling) + 32);
  Bytes occ((size_t)orows * ocols + 64, 0);
  std::vector<ScoredPoint> kept;
  for (const ScoredPoint& p : *points) {
    const int cy = (int)(p.y * scaling + 16), cx = (int)(p.x * scaling + 16);
    const double s0 = (double)occ[(size_t)cy * ocols + cx];
    const float nsc1 = sqrtf(sqrtf(p.score / max_score)) * 255.0f;
    if (nsc1 < s0) continue;
    const float nsc = 0.99f * nsc1;
    for (int y = 0; y < 31; ++y)
      for (int x = 0; x < 31; ++x) {
        uint8_t& o = occ[(size_t)(cy + y - 15) * ocols + cx + x - 15];
        const int add = (int)ceil(lut[y][x] * nsc) & 0xFF;
        o = (uint8_t)std::min(255, (int)o + add);
      }
    kept.push_back(p);
    if (kept.size() == max_kpt) break;
  }
  points->swap(kept);
}

// reference scale-space-feature-detector.h:100-128 + scale-space-layer-inl.h:60-428.
void HarrisDetect(const uint8_t* image, int w, int h, int octaves, double radius, double abs_thr, size_t max_kpt,
                  std::vector<orc_keypoint>* out) {
  out->clear();
  if (w == 0 || h == 0) return;
  const int n = std::max(octaves * 2, 1);
  std::vector<HarrisLayer> L(n);
  for (int i = 0; i < n; ++i) {
    HarrisLayer& l = L[i];
    l.number = i;
    if (i == 0) {
      l.w = w; l.h = h; l.img.assign(image, image + (size_t)w * h); l.octave = true;
    } else if (i == 1) {
      l.w = 2 * (w / 3); l.h = 2 * (h / 3); l.img.resize((size_t)l.w * l.h);
      Twothirdsample8(L[0].img.data(), w, h, l.img.data()); l.octave = false;
    } else {
      const HarrisLayer& s = L[i - 2];
      l.w = s.w / 2; l.h = s.h / 2; l.img.resize((size_t)l.w * l.h);
      Halfsample8(s.img.data(), s.w, s.h, l.img.data()); l.octave = (i % 2 == 0);
    }
    if (l.octave) {
      l.offset_above = -0.25; l.offset_below = 1.0 / 6.0; l.scale_above = 2.0 / 3.0; l.scale_below = 4.0 / 3.0;
      l.scale = i == 0 ? 1.0 : pow(2.0, (double)(i / 2));
    } else {
      l.offset_above = -1.0 / 6.0; l.offset_below = 0.125; l.scale_above = 0.75; l.scale_below = 1.5;
      l.scale = pow(2.0, (double)(i / 2)) * 1.5;
    }
    l.offset = i == 0 ? 0.0 : l.scale * 0.5 - 0.5;
    l.scores.resize((size_t)l.w * l.h);
    HarrisScores(l.img.data(), l.w, l.h, l.scores.data());
  }
  double r = radius;
  if (r == 0) r = 1;  // SetUniformityRadius
  const bool uniformity = radius > 0.0;
  for (int i = 0; i < n; ++i) {
    HarrisLayer& l = L[i];
    std::vector<ScoredPoint> pts;
    HarrisMaxima(l.scores.data(), l.w, l.h, (int)abs_thr, &pts);
    const HarrisLayer* above = i + 1 < n ? &L[i + 1] : nullptr;
    const HarrisLayer* below = i > 0 ? &L[i - 1] : nullptr;
    if (above || below) {
      std::vector<ScoredPoint> kept;
      const int oa = (int)(1.0 / l.scale_above);  // == 1
      const int ob = (int)(1.0 / l.scale_below);  // == 0
      static const int kOff[9][2] = {{0, 0}, {1, 0}, {-1, 0}, {0, 1}, {0, -1}, {1, 1}, {1, -1}, {-1, 1}, {-1, -1}};
      for (const ScoredPoint& p : pts) {
        const int c = p.score;
        if (c < (int)abs_thr) continue;
        bool ok = true;
        if (above)
          for (int k = 0; k < 9 && ok; ++k) {
            const double u = (int)p.x + kOff[k][0] * oa, v = (int)p.y + kOff[k][1] * oa;
            if (c < above->ScoreD(l.scale_above * (u + l.offset_above), l.scale_above * (v + l.offset_above))) ok = false;
          }
        if (below)
          for (int k = 0; k < 9 && ok; ++k) {
            const double u = (int)p.x + kOff[k][0] * ob, v = (int)p.y + kOff[k][1] * ob;
            if (c < below->ScoreD(l.scale_below * (u + l.offset_below), l.scale_below * (v + l.offset_below))) ok = false;
          }
        if (ok) kept.push_back(p);
      }
      pts.swap(kept);
    }
    if (pts.empty()) continue;
    if (uniformity && r > 0.0) EnforceUniformity(r, l.h, l.w, max_kpt, &pts);
    else {
      // KeyPointBucketing (key-point-bucketing-inl.h:74-112) with the 4 x 4 buckets of ScaleSpaceLayer
      // (scale-space-layer.h:73-74): KeyPointBuckets::filterKeyPoints (:44-72) sorts, then keeps a point while its
      // bucket holds fewer than maxNumKpt / 16 points (key-point-bucketing.h:64-66, an unsigned int).  (The
      // single-bucket partial_sort branch needs numBuckets == 1, which ScaleSpaceLayer never sets.  The
      // reference reserves maxNumKpt entries first, so it throws for the default SIZE_MAX; callers pass a limit.)
      const unsigned step_u = 1u + (unsigned)((l.w - 1u) / 4u), step_v = 1u + (unsigned)((l.h - 1u) / 4u);
      const unsigned quota = (unsigned)(max_kpt / 16u);
      unsigned count[4][4] = {};
      std::sort(pts.begin(), pts.end());
      std::vector<ScoredPoint> kept;
      for (const ScoredPoint& p : pts) {
        unsigned* c = &count[p.x / step_u][p.y / step_v];
        if (*c < quota) { ++*c; kept.push_back(p); }
      }
      pts.swap(kept);
    }
    for (const ScoredPoint& p : pts) {
      const int u = p.x, v = p.y;
      auto S = [&](int uu, int vv) { return (double)l.scores[(size_t)vv * l.w + uu]; };
      float dx, dy;
      HarrisSubpixel2D(S(u - 1, v - 1), S(u, v - 1), S(u + 1, v - 1), S(u - 1, v), S(u, v), S(u + 1, v), S(u - 1, v + 1),
                       S(u, v + 1), S(u + 1, v + 1), &dx, &dy);
      orc_keypoint k;
      k.x = (float)(l.scale * ((p.x + dx) + l.offset));
      k.y = (float)(l.scale * ((p.y + dy) + l.offset));
      k.size = (float)(l.scale * 12.0);
      k.angle = -1; k.response = (float)p.score; k.octave = l.number / 2; k.class_id = -1;
      out->push_back(k);
    }
  }
}

// reference scale-space-feature-detector.h:100-128 with a NON-EMPTY key-point vector ("use passed key points",
// :103-108) and scale-space-layer-inl.h:193-208,370-428: no scores are computed; the points with response > 1e6
// become (int score, uint16 x, uint16 y), go through the uniformity enforcement (or the bucketing) and come back
// unrefined.  Only one layer (octaves == 0): a second layer would index its smaller occupancy map with layer 0's
// image coordinates (out of bounds in the reference).  If no point passes the response test the vector is
// returned untouched (:370-371 returns before the clear).
void HarrisFilterPassed(int w, int h, double radius, size_t max_kpt, const orc_keypoint* in, int n_in,
                        std::vector<orc_keypoint>* out) {
  std::vector<ScoredPoint> pts;
  for (int k = 0; k < n_in; ++k)
    if (in[k].response > 1e6) { ScoredPoint p; p.score = (int)in[k].response; p.x = (uint16_t)in[k].x; p.y = (uint16_t)in[k].y; pts.push_back(p); }
  out->clear();
  if (pts.empty()) { out->assign(in, in + n_in); return; }
  const double r = radius == 0 ? 1.0 : radius;  // SetUniformityRadius (scale-space-layer-inl.h:185-189)
  if (radius > 0.0 && r > 0.0) EnforceUniformity(r, h, w, max_kpt, &pts);
  else {
    const unsigned step_u = 1u + (unsigned)((w - 1u) / 4u), step_v = 1u + (unsigned)((h - 1u) / 4u);
    const unsigned quota = (unsigned)(max_kpt / 16u);
    unsigned count[4][4] = {};
    std::sort(pts.begin(), pts.end());
    std::vector<ScoredPoint> kept;
    for (const ScoredPoint& p : pts) {
      unsigned* c = &count[p.x / step_u][p.y / step_v];
      if (*c < quota) { ++*c; kept.push_back(p); }
    }
    pts.swap(kept);
  }
  for (const ScoredPoint& p : pts) {
    orc_keypoint k;
    k.x = (float)(1.0 * ((p.x) + 0.0)); k.y = (float)(1.0 * ((p.y) + 0.0));
    k.size = (float)(1.0 * 12.0); k.angle = -1; k.response = (float)p.score; k.octave = 0; k.class_id = -1;
    out->push_back(k);
  }
}

}  // namespace

extern "C" {

void orc_halfsample8(const uint8_t* src, int w, int h, uint8_t* dst) { Halfsample8(src, w, h, dst); }
void orc_twothirdsample8(const uint8_t* src, int w, int h, uint8_t* dst) { Twothirdsample8(src, w, h, dst); }
void orc_thrmap(const uint8_t* img, int w, int h, uint8_t* thr) { ThresholdMap(img, w, h, thr); }

void orc_dense_scores(const uint8_t* img, int w, int h, uint8_t* out916, uint8_t* out58) {
  Layer l; l.w = w; l.h = h; l.img.assign(img, img + (size_t)w * h); l.scores.assign((size_t)w * h, 0);
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      if (out916) out916[(size_t)y * w + x] = (uint8_t)l.Score(x, y, 1);
      if (out58) out58[(size_t)y * w + x] = (uint8_t)l.Score58(x, y, 1);
    }
}

int orc_layer_corners(const uint8_t* img, int w, int h, int thresh, int lower, int32_t* xys, int cap) {
  Bytes thr((size_t)w * h);
  ThresholdMap(img, w, h, thr.data());
  std::vector<Corner> c;
  DetectCorners(img, thr.data(), w, h, thresh, lower, 230, &c);
  for (size_t i = 0; i < c.size() && (int)i < cap; ++i) { xys[3 * i] = c[i].x; xys[3 * i + 1] = c[i].y; xys[3 * i + 2] = c[i].score; }
  return (int)c.size();
}

int orc_pyramid(const uint8_t* img, int w, int h, int octaves, uint8_t* out, int32_t* dims, float* scale_offset) {
  ScaleSpace ss;
  ss.Construct(img, w, h, octaves, 60);
  size_t off = 0;
  for (size_t i = 0; i < ss.L.size(); ++i) {
    dims[2 * i] = ss.L[i].w; dims[2 * i + 1] = ss.L[i].h;
    if (scale_offset) { scale_offset[2 * i] = ss.L[i].scale; scale_offset[2 * i + 1] = ss.L[i].offset; }
    if (out) memcpy(out + off, ss.L[i].img.data(), ss.L[i].img.size());
    off += ss.L[i].img.size();
  }
  return (int)ss.L.size();
}

int orc_agast_detect(const uint8_t* img, int w, int h, int thresh, int octaves, int suppress, const uint8_t* mask,
                     orc_keypoint* out, int cap) {
  ScaleSpace ss;
  ss.suppress = suppress != 0;
  ss.Construct(img, w, h, octaves, thresh);
  std::vector<orc_keypoint> kps;
  ss.GetKeypoints(&kps);
  // reference brisk-feature-detector.cc:49-66 (RemoveInvalidKeyPoints)
  int n = 0, total = 0;
  for (const orc_keypoint& k : kps) {
    if (mask && mask[(size_t)(int)(k.y + 0.5f) * w + (int)(k.x + 0.5f)] == 0) continue;
    if (n < cap) out[n++] = k;
    ++total;
  }
  return total;
}

// reference brisk-feature-detector.cc:87-92 (BriskFeatureDetector::ComputeScale) with a non-empty key-point
// vector.  suppress = 0 is only defined for one layer (as in orc_agast_detect).
int orc_compute_scale(const uint8_t* img, int w, int h, int thresh, int octaves, int suppress, const orc_keypoint* in,
                      int n_in, orc_keypoint* out, int cap) {
  if (n_in <= 0) return -1;
  if (!suppress && octaves != 0) return -1;
  ScaleSpace ss;
  ss.suppress = suppress != 0;
  ss.Construct(img, w, h, octaves, thresh);
  std::vector<orc_keypoint> kps;
  ss.GetKeypointsProvided(in, n_in, &kps);
  for (size_t i = 0; i < kps.size() && (int)i < cap; ++i) out[i] = kps[i];
  return (int)kps.size();
}

// debug aid: final lazy-cache state of all layers, concatenated.
int orc_agast_cache_dump(const uint8_t* img, int w, int h, int thresh, int octaves, uint8_t* out) {
  ScaleSpace ss;
  ss.Construct(img, w, h, octaves, thresh);
  std::vector<orc_keypoint> kps;
  ss.GetKeypoints(&kps);
  size_t off = 0;
  for (auto& l : ss.L) { memcpy(out + off, l.scores.data(), l.scores.size()); off += l.scores.size(); }
  return (int)kps.size();
}

void orc_integral8(const uint8_t* img, int w, int h, int32_t* out) { Integral8(img, w, h, out); }

int orc_describe(const uint8_t* img, int w, int h, orc_keypoint* kps, int n, int rot, int scale, int version,
                 float pattern_scale, uint8_t* desc, int32_t* desc_bytes) {
  std::shared_ptr<Pattern> p = GetPattern(version, pattern_scale);
  if (desc_bytes) *desc_bytes = p->strings;
  return Describe(img, w, h, kps, n, rot != 0, scale != 0, *p, desc);
}

int orc_pattern_dump(int version, float pattern_scale, int32_t* counts, float* points_xys, float* scale_list,
                     uint32_t* size_list, uint32_t* short_pairs, int32_t* long_pairs) {
  std::shared_ptr<Pattern> p = GetPattern(version, pattern_scale);
  if (counts) { counts[0] = p->points; counts[1] = (int)p->shorts.size(); counts[2] = (int)p->longs.size(); counts[3] = p->strings; }
  if (points_xys) memcpy(points_xys, p->pts.data(), sizeof(PatternPoint) * p->pts.size());
  if (scale_list) memcpy(scale_list, p->scale_list, sizeof(p->scale_list));
  if (size_list) memcpy(size_list, p->size_list, sizeof(p->size_list));
  if (short_pairs) for (size_t i = 0; i < p->shorts.size(); ++i) { short_pairs[2 * i] = p->shorts[i].i; short_pairs[2 * i + 1] = p->shorts[i].j; }
  if (long_pairs)
    for (size_t i = 0; i < p->longs.size(); ++i) {
      long_pairs[4 * i] = p->longs[i].i; long_pairs[4 * i + 1] = p->longs[i].j;
      long_pairs[4 * i + 2] = p->longs[i].wdx; long_pairs[4 * i + 3] = p->longs[i].wdy;
    }
  return 0;
}

void orc_harris_scores(const uint8_t* img, int w, int h, int32_t* out) { HarrisScores(img, w, h, out); }

int orc_harris_maxima(const uint8_t* img, int w, int h, int abs_thr, int32_t* out_sxy, int cap) {
  std::vector<int32_t> sc((size_t)w * h);
  HarrisScores(img, w, h, sc.data());
  std::vector<ScoredPoint> pts;
  HarrisMaxima(sc.data(), w, h, abs_thr, &pts);
  for (size_t i = 0; i < pts.size() && (int)i < cap; ++i) { out_sxy[3 * i] = pts[i].score; out_sxy[3 * i + 1] = pts[i].x; out_sxy[3 * i + 2] = pts[i].y; }
  return (int)pts.size();
}

int orc_harris_detect(const uint8_t* img, int w, int h, int octaves, double radius, double abs_thr, int64_t max_kpt,
                      orc_keypoint* out, int cap) {
  std::vector<orc_keypoint> kps;
  HarrisDetect(img, w, h, octaves, radius, abs_thr, max_kpt < 0 ? (size_t)-1 : (size_t)max_kpt, &kps);
  for (size_t i = 0; i < kps.size() && (int)i < cap; ++i) out[i] = kps[i];
  return (int)kps.size();
}

int orc_harris_detect_passed(int w, int h, double radius, int64_t max_kpt, const orc_keypoint* in, int n_in,
                             orc_keypoint* out, int cap) {
  if (n_in <= 0 || w <= 0 || h <= 0) return -1;
  std::vector<orc_keypoint> kps;
  HarrisFilterPassed(w, h, radius, max_kpt < 0 ? SIZE_MAX : (size_t)max_kpt, in, n_in, &kps);
  for (size_t i = 0; i < kps.size() && (int)i < cap; ++i) out[i] = kps[i];
  return (int)kps.size();
}

// reference hamming-inl.h:85-134: popcount of the XOR over 16-byte words.
int orc_hamming(const uint8_t* a, const uint8_t* b, int nbytes) {
  int r = 0;
  for (int i = 0; i < nbytes; ++i) r += __builtin_popcount((unsigned)(a[i] ^ b[i]));
  return r;
}

// reference brute-force-matcher.cc:80-162 (commonKnnMatchImpl), one train
// image, no mask: k successive arg-mins, first minimum wins (lowest train
// index on ties), picked entries overwritten with INT_MAX; output sorted by
// distance (already is).
int orc_knn(const uint8_t* q, int64_t nq, const uint8_t* t, int64_t nt, int nbytes, int k, int32_t* idx, int32_t* dist) {
  std::vector<int> d(nt);
  for (int64_t i = 0; i < nq; ++i) {
    for (int64_t j = 0; j < nt; ++j) d[j] = orc_hamming(q + i * nbytes, t + j * nbytes, nbytes);
    for (int kk = 0; kk < k; ++kk) {
      int best = 0x7fffffff; int64_t bj = -1;
      for (int64_t j = 0; j < nt; ++j) if (d[j] < best) { best = d[j]; bj = j; }
      idx[i * k + kk] = (int32_t)bj; dist[i * k + kk] = bj < 0 ? -1 : best;
      if (bj >= 0) d[bj] = 0x7fffffff;
    }
  }
  return 0;
}

// std::sort(curMatches->begin(), curMatches->end()) of brute-force-matcher.cc:160,210 on cv::DMatch
// records (operator< compares `distance` only): libstdc++'s introsort, not stable, so the permutation of
// equal distances is part of the reference's result.  In/out: parallel arrays of n matches.
int orc_sort_matches(int32_t* train_idx, int32_t* img_idx, float* distance, int n) {
  struct M { int t, i; float d; bool operator<(const M& m) const { return d < m.d; } };
  std::vector<M> v(n);
  for (int j = 0; j < n; ++j) v[j] = M{train_idx[j], img_idx[j], distance[j]};
  std::sort(v.begin(), v.end());
  for (int j = 0; j < n; ++j) { train_idx[j] = v[j].t; img_idx[j] = v[j].i; distance[j] = v[j].d; }
  return 0;
}

// std::sort(keypoints.begin(), keypoints.end(), compareKeypointScore) of harris-feature-detector.cc:50-52,306: descending by the
// float response, libstdc++'s introsort (the permutation of equal responses is part of the result).  perm: in/out indices.
int orc_sort_desc_by_response(const float* response, int32_t* perm, int n) {
  struct K { float r; int32_t i; };
  std::vector<K> v(n);
  for (int j = 0; j < n; ++j) v[j] = K{response[perm[j]], perm[j]};
  std::sort(v.begin(), v.end(), [](K a, K b) { return a.r > b.r; });
  for (int j = 0; j < n; ++j) perm[j] = v[j].i;
  return 0;
}

// All pairwise distances (hamming-inl.h:85-134) of two descriptor sets: out[nq][nt].
int orc_hamming_matrix(const uint8_t* q, int64_t nq, const uint8_t* t, int64_t nt, int nbytes, int32_t* out) {
  for (int64_t i = 0; i < nq; ++i)
    for (int64_t j = 0; j < nt; ++j) out[i * nt + j] = orc_hamming(q + i * nbytes, t + j * nbytes, nbytes);
  return 0;
}

}  // extern "C"
