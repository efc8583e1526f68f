// Per-sample arithmetic of the BRISK descriptor (reference
// brisk/src/brisk-descriptor-extractor.cc:370-530,612-778), as __host__
// __device__ functions shared by the describe kernel and the host-side tests.
// Float/double promotions follow the reference expression by expression
// (SURVEY.md App. A.6); compile without FMA contraction.
#pragma once
#include <math.h>

#include "brisk_common.cuh"

namespace briskb200 {

// SmoothedIntensity<uchar,int> (brisk-descriptor-extractor.cc:370-530): box
// filtered intensity (x ~1024) of the pattern point (px, py, sigma) placed at
// key point (kx, ky).  `integral` is the (h+1)x(w+1) int32 integral image with
// row stride iw = w + 1; the image has row pitch `pitch`.
// The two integer normalisation constants of a pattern point depend on its sigma only (:387,
// :412-413); they are tabulated per (scale, point) on the host with this function so that the kernel
// does no double-precision division.
BRISK_HD void sampling_constants(float sigma_half, int* scaling, int* scaling2) {
  const float area = (float)(4.0 * (double)sigma_half * (double)sigma_half);
  *scaling = (int)(4194304.0 / (double)area);
  *scaling2 = (int)((double)((float)*scaling * area) / 1024.0);
}

// The twelve integral-image taps of the large-box case (:459-484) and the four weighted corner pixels the reference's
// pointer walk reads (:447-456), by name (t1..t12 as in the reference).
struct BoxTaps {
  int t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, t11, t12;
  int p_tl, p_tr;  // I(y_top, x_left), I(y_top, x_left + dx + 1)
  int p_bl, p_br;  // I(y_top + dy, x_left + 1), I(y_top + dy, x_left + dx + 2): one row up and one column right of the
                   // geometric bottom corners -- what the reference reads
};

// Plain (h+1) x (w+1) int32 integral image, row stride iw = w + 1: one load per tap.
struct PlainIntegral {
  const int32_t* s;
  int iw;
  BRISK_HD void taps(const uint8_t* p /* &img(y_top, x_left) */, int pitch, int y_top, int x_left, int dx, int dy, BoxTaps* o) const {
    const int32_t* q = s + (long long)y_top * iw + x_left + 1;
    o->t1 = q[0]; o->t2 = q[dx];
    const int32_t* q1 = q + iw;
    o->t3 = q1[dx]; o->t4 = q1[dx + 1]; o->t12 = q1[0]; o->t11 = q1[-1];
    const int32_t* q2 = q1 + (long long)dy * iw;
    o->t5 = q2[dx + 1]; o->t6 = q2[dx]; o->t9 = q2[0]; o->t10 = q2[-1];
    const int32_t* q3 = q2 + iw;
    o->t7 = q3[dx]; o->t8 = q3[0];
    o->p_tl = p[0]; o->p_tr = p[dx + 1];
    const uint8_t* pc = p + (dx + 1) + (long long)dy * pitch + 1;
    o->p_br = pc[0]; o->p_bl = pc[-(dx + 1)];
  }
};

// The same integral image stored as one 16-byte BLOCK per pixel (Y < h, X < w, row stride w blocks) that holds the 2x2
// neighbourhood a = S(Y,X), b = S(Y,X+1), c = S(Y+1,X), d = S(Y+1,X+1) and two pixels:
//   word 0 = a,  word 1 = (b - a) | I(Y-1, X+1) << 24,  word 2 = (c - a) | I(Y, X) << 24,  word 3 = d.
// b - a is a column sum and c - a a row sum (< 2^24 for images up to 65 793 pixels a side).  The twelve taps are three
// corners each of the four blocks at rows {y_top, y_top+dy+1} x columns {x_left, x_left+dx+1}; the top blocks carry the
// box's top corner pixels (their own), the bottom blocks the two pixels the reference reads for the bottom corners
// (one row up, one column right): four 16-byte loads replace twelve 4-byte and four 1-byte loads.  I(Y-1, X+1) is taken
// from the tightly packed image the reference samples (stride == width), so for X + 1 == w it is I(Y, 0).
struct Block4 { int a, b, c, d; };   // as stored (encoded)
struct BlockFields { int a, b, c, d, pix, up_right; };
BRISK_HD Block4 encode_block(int a, int b, int c, int d, int up_right) {
  const unsigned pix = (unsigned)(d - b - c + a);
  return Block4{a, (int)(((unsigned)(b - a) & 0xffffffu) | ((unsigned)up_right << 24)), (int)(((unsigned)(c - a) & 0xffffffu) | (pix << 24)), d};
}
BRISK_HD BlockFields decode_block(const Block4& e) {
  BlockFields f;
  f.a = e.a;
  f.b = (int)((unsigned)e.a + ((unsigned)e.b & 0xffffffu));
  f.c = (int)((unsigned)e.a + ((unsigned)e.c & 0xffffffu));
  f.d = e.d;
  f.pix = (int)((unsigned)e.c >> 24);
  f.up_right = (int)((unsigned)e.b >> 24);
  return f;
}
struct BlockIntegral {
  const Block4* s;
  int bw;  // blocks per row (= image width)
  BRISK_HD static BlockFields load(const Block4* p) {
#ifdef __CUDA_ARCH__
    const int4 v = __ldg(reinterpret_cast<const int4*>(p));   // (L1-allocating on purpose: the boxes of neighbouring pattern points overlap)
    return decode_block(Block4{v.x, v.y, v.z, v.w});
#else
    return decode_block(*p);
#endif
  }
  BRISK_HD void taps(const uint8_t*, int, int y_top, int x_left, int dx, int dy, BoxTaps* o) const {
    const Block4* r0 = s + (long long)y_top * bw + x_left;
    const Block4* r1 = r0 + (long long)(dy + 1) * bw;
    const BlockFields tl = load(r0), tr = load(r0 + dx + 1), bl = load(r1), br = load(r1 + dx + 1);
    o->t1 = tl.b; o->t11 = tl.c; o->t12 = tl.d;
    o->t2 = tr.a; o->t3 = tr.c; o->t4 = tr.d;
    o->t10 = bl.a; o->t9 = bl.b; o->t8 = bl.d;
    o->t6 = br.a; o->t5 = br.b; o->t7 = br.c;
    o->p_tl = tl.pix; o->p_tr = tr.pix;
    o->p_bl = bl.up_right; o->p_br = br.up_right;
  }
};

template <class Integral>
BRISK_HD int smoothed_intensity_t(const uint8_t* __restrict__ img, int pitch, const Integral& integral,
                                  float kx, float ky, float px, float py, float sigma_half, int scaling, int scaling2) {
  const float xf = px + kx, yf = py + ky;
  if ((double)sigma_half < 0.5) {
    const int x = (int)xf, y = (int)yf;
    const int r_x = (int)((xf - (float)x) * 1024.0f), r_y = (int)((yf - (float)y) * 1024.0f);
    const int r_x_1 = 1024 - r_x, r_y_1 = 1024 - r_y;
    const uint8_t* p = img + (long long)y * pitch + x;
    int v = r_x_1 * r_y_1 * (int)p[0];
    v += r_x * r_y_1 * (int)p[1];
    v += r_x * r_y * (int)p[pitch + 1];
    v += r_x_1 * r_y * (int)p[pitch];
    return v / 1024;
  }
  const float x_1 = xf - sigma_half, x1 = xf + sigma_half, y_1 = yf - sigma_half, y1 = yf + sigma_half;
  const int x_left = (int)((double)x_1 + 0.5), y_top = (int)((double)y_1 + 0.5);
  const int x_right = (int)((double)x1 + 0.5), y_bottom = (int)((double)y1 + 0.5);
  const float r_x_1 = (float)((double)((float)x_left - x_1) + 0.5), r_y_1 = (float)((double)((float)y_top - y_1) + 0.5);
  const float r_x1 = (float)((double)(x1 - (float)x_right) + 0.5), r_y1 = (float)((double)(y1 - (float)y_bottom) + 0.5);
  const int dx = x_right - x_left - 1, dy = y_bottom - y_top - 1;
  const float fs = (float)scaling;
  const int A = (int)((r_x_1 * r_y_1) * fs), B = (int)((r_x1 * r_y_1) * fs);
  const int C = (int)((r_x1 * r_y1) * fs), D = (int)((r_x_1 * r_y1) * fs);
  const int r_x_1_i = (int)(r_x_1 * fs), r_y_1_i = (int)(r_y_1 * fs), r_x1_i = (int)(r_x1 * fs), r_y1_i = (int)(r_y1 * fs);
  const uint8_t* p = img + (long long)y_top * pitch + x_left;
  if (dx + dy > 2) {
    BoxTaps t;
    integral.taps(p, pitch, y_top, x_left, dx, dy, &t);
    // four weighted corner pixels; the reference's pointer walk (:447-456)
    // reads the bottom pair one row up and one column right of the geometric
    // corners -- reproduced.
    int v = A * t.p_tl + B * t.p_tr;
    v += C * t.p_br + D * t.p_bl;
    // twelve integral-image taps (:459-484)
    const int upper = (t.t3 - t.t2 + t.t1 - t.t12) * r_y_1_i;
    const int middle = (t.t6 - t.t3 + t.t12 - t.t9) * scaling;
    const int left = (t.t9 - t.t12 + t.t11 - t.t10) * r_x_1_i;
    const int right = (t.t5 - t.t4 + t.t3 - t.t6) * r_x1_i;
    const int bottom = (t.t7 - t.t6 + t.t9 - t.t8) * r_y1_i;
    return (v + upper + middle + left + right + bottom) / scaling2;
  }
  // small box: weighted sum over the (dx+2) x (dy+2) window, written as the
  // reference's index walk (:497-529) so that the degenerate windows that float
  // rounding can produce when 2*sigma ~ 1 (dx or dy == -1: the walk then reuses /
  // drifts over neighbouring pixels) come out the same.
  long long i = 0;
  int v = A * (int)p[i];
  ++i;
  for (const long long e = i + dx; i < e; ++i) v += r_y_1_i * (int)p[i];
  v += B * (int)p[i];
  i += pitch - dx - 1;
  for (const long long ej = i + (long long)dy * pitch; i < ej; i += pitch - dx - 1) {
    v += r_x_1_i * (int)p[i];
    ++i;
    for (const long long e = i + dx; i < e; ++i) v += (int)p[i] * scaling;
    v += r_x1_i * (int)p[i];
  }
  v += D * (int)p[i];
  ++i;
  for (const long long e = i + dx; i < e; ++i) v += r_y1_i * (int)p[i];
  v += C * (int)p[i];
  return v / scaling2;
}

BRISK_HD int smoothed_intensity(const uint8_t* __restrict__ img, int pitch, const int32_t* __restrict__ integral, int iw,
                                float kx, float ky, float px, float py, float sigma_half, int scaling, int scaling2) {
  return smoothed_intensity_t(img, pitch, PlainIntegral{integral, iw}, kx, ky, px, py, sigma_half, scaling, scaling2);
}

// Rotation bin of a freshly estimated orientation (:732-739): angle in degrees
// from the long-pair gradient, theta = its 1024-bin index.
BRISK_HD float orientation_angle(int d0, int d1) {
  return (float)(atan2((double)(float)d1, (double)(float)d0) / 3.14159265358979323846 * 180.0);
}
BRISK_HD int theta_from_estimated(float angle) {
  int theta = (int)((double)(1024.0f * angle) / 360.0 + 0.5);
  if (theta < 0) theta += 1024;
  if (theta >= 1024) theta -= 1024;
  return theta;
}
// Rotation bin of a caller-supplied angle (:746-751).  The reference wraps once,
// which covers angles in (-360, 720) degrees; beyond that it indexes its table
// out of bounds (undefined).  Such angles are reduced modulo one turn here.
BRISK_HD int theta_from_given(float angle) {
  int theta = (int)(1024.0 * ((double)angle / 360.0) + 0.5);
  if (theta < 0) theta += 1024;
  if (theta >= 1024) theta -= 1024;
  if (theta < 0 || theta >= 1024) theta = ((theta % 1024) + 1024) % 1024;
  return theta;
}

}  // namespace briskb200
