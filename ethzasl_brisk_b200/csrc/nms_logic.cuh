// Scale-space non-maximum suppression and refinement of the AGAST path
// (reference brisk/src/brisk-scale-space.cc:92-1099), reformulated so that it
// can run data-parallel.
//
// The reference evaluates FAST scores lazily and caches them in a byte map
// (brisk-layer.cc:118-132); a cached value > 2 is returned whatever threshold
// is asked, so its results depend on the order in which pixels were first
// looked at.  Observations that remove the sequential replay:
//
//  * Every look-up with threshold 1 returns a pure function of the image:
//    T(q) at a detected corner, else F(q) = fast916 score if F >= 1, else 0.
//    All of Refine3D / GetScoreMaxAbove / GetScoreMaxBelow use threshold 1, so
//    their VALUES are order independent; only their cache FOOTPRINT matters.
//  * IsMax2D's eight comparisons `center < s` are order independent as well
//    (a cached neighbour returns F, an uncached one F or 0, and both compare the
//    same against center).  What depends on the cache is only the tie path: the
//    smoothed-centre comparison reads raw cache bytes.
//  * The cache state of pixel q of layer i at the time corner c is examined is
//    a function of (a) whether the layer below touched q (all of that happens
//    before layer i is processed), and (b) the look-ups made by raster-earlier
//    corners of layer i within 2 pixels of q: their IsMax2D neighbour look-ups
//    (threshold = their own score) and, if they were accepted, their
//    threshold-1 patch look-ups.
//
// So: `nms_prefix` (parallel) runs the eight comparisons for every corner,
// `nms_checks` (parallel) runs the scale-space checks without side effects,
// `nms_tie_decide` resolves the tying corners of a layer in raster order using
// only byte look-ups, `mark_above` (parallel) records the cache footprint a
// layer leaves on the layer above, and `refine_emit` (parallel) produces the
// key points.  All functions are __host__ __device__ so that tests can run the
// very same code on the CPU against the oracle.
#pragma once
#include "brisk_math.cuh"
#include "fast_packed.cuh"

namespace briskb200 {

struct LayerView {
  const uint8_t* img;  // layer image
  uint16_t* cm;        // corner map (see brisk_common.cuh)
  uint8_t* bm;         // 1 where the layer below looked the pixel up (threshold 1)
  int w, h, pitch;
  float scale, offset;
};

enum LayerMode { kModeMid = 0, kModeLast = 1, kModeSingle = 2 };

// Neighbour order of IsMax2D's look-ups (brisk-scale-space.cc:439-461).
BRISK_HD void isMax2dOffset(int j, int* dx, int* dy) {
  // (-1,0) (1,0) (0,-1) (0,1) (-1,1) (1,1) (1,-1) (-1,-1)
  switch (j) {
    case 0: *dx = -1; *dy = 0; break;
    case 1: *dx = 1; *dy = 0; break;
    case 2: *dx = 0; *dy = -1; break;
    case 3: *dx = 0; *dy = 1; break;
    case 4: *dx = -1; *dy = 1; break;
    case 5: *dx = 1; *dy = 1; break;
    case 6: *dx = 1; *dy = -1; break;
    default: *dx = -1; *dy = -1; break;
  }
}

// Index in that order of the neighbour at offset (dx, dy) in [-1,1]^2 \ (0,0).
BRISK_HD int isMax2dIndex(int dx, int dy) {
  // row dy=-1: (-1,-1)->7 (0,-1)->2 (1,-1)->6 ; dy=0: (-1,0)->0 (1,0)->1 ; dy=1: (-1,1)->4 (0,1)->3 (1,1)->5
  const int code = (dy + 1) * 3 + (dx + 1);
  switch (code) {
    case 0: return 7;
    case 1: return 2;
    case 2: return 6;
    case 3: return 0;
    case 5: return 1;
    case 6: return 4;
    case 7: return 3;
    default: return 5;
  }
}

BRISK_HD bool in_border(const LayerView& L, int x, int y) { return x < 3 || y < 3 || x >= L.w - 3 || y >= L.h - 3; }

// FAST score clipped at 0 (values <= 0 are indistinguishable for every use).
BRISK_HD int fastF(const LayerView& L, int x, int y) {
  const int f = fast916(L.img, L.pitch, x, y);
  return f < 0 ? 0 : f;
}

// Value of BriskLayer::GetAgastScore(x, y, 1) (brisk-layer.cc:118-132): pure.
BRISK_HD int score1(const LayerView& L, int x, int y) {
  if (in_border(L, x, y)) return 0;
  const int t = L.cm[(long long)y * L.pitch + x] & kCmT;
  if (t) return t;
  return fastF(L, x, y);  // F >= 1 ? F : 0
}

// BriskLayer::GetAgastScore_5_8(x, y, 1) (brisk-layer.cc:134-145).
BRISK_HD int score58(const LayerView& L, int x, int y) {
  if (x < 2 || y < 2 || x >= L.w - 2 || y >= L.h - 2) return 0;
  const int f = fast58(L.img, L.pitch, x, y);
  return f < 1 ? 0 : f;
}

// Scores of a small pixel rectangle (at most 4x4) of one layer, evaluated once
// and kept in two registers.  The scans of the neighbouring layers touch the
// same few pixels over and over (every bilinear read covers a 2x2 cell), so all
// their look-ups go through a tile instead of re-running the FAST arithmetic.
struct ScoreTile {
  unsigned long long lo, hi;  // 16 bytes, row stride 4
  int x0, y0;
#ifdef BRISK_TILE_CHECK
  int x1, y1;
  mutable int violations;
#endif
  BRISK_HD int get(int x, int y) const {
#ifdef BRISK_TILE_CHECK
    if (x < x0 || y < y0 || x > x1 || y > y1) { ++violations; return 0; }
#endif
    const int i = ((y - y0) << 2) + (x - x0);
    const unsigned long long w = i < 8 ? lo : hi;
    return (int)((w >> ((i & 7) << 3)) & 0xffull);
  }
};

// Tile over [xa, xb] x [ya, yb] (inclusive; at most 4x4).
BRISK_HD void fill_tile(const LayerView& L, int xa, int ya, int xb, int yb, ScoreTile* t) {
  t->lo = 0; t->hi = 0; t->x0 = xa; t->y0 = ya;
#ifdef BRISK_TILE_CHECK
  t->x1 = xb; t->y1 = yb; t->violations = (xb - xa > 3 || yb - ya > 3) ? 1 : 0;
#endif
#ifdef __CUDA_ARCH__
  // device: one packed evaluation per tile row (four pixels at once, fast_packed.cuh); corners keep their T
#pragma unroll 1
  for (int y = ya; y <= yb; ++y) {
    uint32_t f[2];
    fast916_row<2>(L.img, L.pitch, L.h, xa, y, f);
    unsigned long long rowbits = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int x = xa + j;
      if (x > xb || in_border(L, x, y)) continue;
      const int tt = L.cm[(long long)y * L.pitch + x] & kCmT;
      const int v = tt ? tt : (int)((f[j >> 1] >> ((j & 1) * 16)) & 0xffu);
      rowbits |= (unsigned long long)v << (8 * j);
    }
    const int i = (y - ya) << 2;
    if (i < 8) t->lo |= rowbits << (i << 3); else t->hi |= rowbits << ((i & 7) << 3);
  }
#else
  for (int y = ya; y <= yb; ++y)
    for (int x = xa; x <= xb; ++x) {
      const int i = ((y - ya) << 2) + (x - xa);
      const unsigned long long v = (unsigned long long)score1(L, x, y);
      if (i < 8) t->lo |= v << (i << 3); else t->hi |= v << ((i & 7) << 3);
    }
#endif
}

// BriskLayer::GetAgastScore(float, float, 1) (brisk-layer.cc:147-161): bilinear
// interpolation of four scores in float, truncated to a byte.
BRISK_HD int tile_score_f(const ScoreTile& t, float xf, float yf) {
  const int x = (int)xf;
  const float rx1 = xf - (float)x;
  const float rx = 1.0f - rx1;
  const int y = (int)yf;
  const float ry1 = yf - (float)y;
  const float ry = 1.0f - ry1;
  const int s00 = t.get(x, y), s10 = t.get(x + 1, y), s01 = t.get(x, y + 1), s11 = t.get(x + 1, y + 1);
  const float v = ((((rx * ry) * (float)s00 + (rx1 * ry) * (float)s10) + (rx * ry1) * (float)s01) + (rx1 * ry1) * (float)s11);
  return (int)(uint8_t)v;
}

BRISK_HD float tile_patch3x3(const ScoreTile& t, int x, int y, float* dx, float* dy) {
  const int s00 = t.get(x - 1, y - 1), s10 = t.get(x, y - 1), s20 = t.get(x + 1, y - 1);
  const int s21 = t.get(x + 1, y), s11 = t.get(x, y), s01 = t.get(x - 1, y);
  const int s02 = t.get(x - 1, y + 1), s12 = t.get(x, y + 1), s22 = t.get(x + 1, y + 1);
  return subpixel2d(s00, s01, s02, s10, s11, s12, s20, s21, s22, dx, dy);
}

BRISK_HD float patch3x3(const LayerView& L, int x, int y, float* dx, float* dy, int* center) {
  int s[9];  // s[3 * column + row], the reference's s_<column>_<row>; one copy of the score code
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
  for (int k = 0; k < 9; ++k) s[k] = score1(L, x + k / 3 - 1, y + k % 3 - 1);
  if (center) *center = s[4];
  return subpixel2d(s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7], s[8], dx, dy);
}

// Pixel rectangle of the neighbouring layer that a scan of the patch
// [x_1,x1]x[y_1,y1] can look up: the 2x2 cells of its float positions, and the
// 3x3 neighbourhoods of its interior positions / its arg-max.
BRISK_HD void scan_tile_bounds(float x_1, float x1, float y_1, float y1, int* xa, int* ya, int* xb_, int* yb_) {
  const int xb = (int)(x_1 + 1), xe = (int)x1, yb = (int)(y_1 + 1), ye = (int)y1;
  // the arg-max can sit at xb or xe even when the interior is empty (xe < xb); its 3x3 patch is read too
  *xa = imin((int)x_1, imin(xb, xe) - 1); *ya = imin((int)y_1, imin(yb, ye) - 1);
  *xb_ = imax((int)x1 + 1, imax(xe, xb) + 1); *yb_ = imax((int)y1 + 1, imax(ye, yb) + 1);
}
BRISK_HD void fill_scan_tile(const LayerView& nb, float x_1, float x1, float y_1, float y1, ScoreTile* t) {
  int xa, ya, xb, yb;
  scan_tile_bounds(x_1, x1, y_1, y1, &xa, &ya, &xb, &yb);
  fill_tile(nb, xa, ya, xb, yb, t);
}
// The same tile from 16 score bytes computed beforehand (row stride 4, origin = the tile's first pixel; nms.cu's
// nms_windows_kernel evaluates them with one lane per pixel).
BRISK_HD void load_scan_tile(const unsigned long long* pre, float x_1, float x1, float y_1, float y1, ScoreTile* t) {
  int xa, ya, xb, yb;
  scan_tile_bounds(x_1, x1, y_1, y1, &xa, &ya, &xb, &yb);
  t->lo = pre[0]; t->hi = pre[1]; t->x0 = xa; t->y0 = ya;
#ifdef BRISK_TILE_CHECK
  t->x1 = xb; t->y1 = yb; t->violations = (xb - xa > 3 || yb - ya > 3) ? 1 : 0;
#endif
}

// Shared scan of GetScoreMaxAbove (brisk-scale-space.cc:757-863) and
// GetScoreMaxBelow (:917-1047) over the patch [x_1,x1]x[y_1,y1] of the
// neighbouring layer (scores served by `nb`).  Returns false when a score above
// `threshold` is met (not tested on the bottom row, as in the reference).  BELOW
// adds the tie rule of :987-1010 on interior pixels.  `steps` receives the
// number of patch positions that were evaluated: the visiting order is fixed,
// so this number identifies the scan's cache footprint (replay_scan_marks).
BRISK_HD bool scan_patch(bool BELOW, const ScoreTile& nb, float x_1, float x1, float y_1, float y1, int threshold, float* max_out,
                         int* mx, int* my, int* steps) {
  // The reference walks a grid of positions in row-major order: columns x_1, the integers xb..xe, x1 and
  // rows y_1, yb..ye, y1 (float positions are bilinear reads, interior ones plain look-ups).  One loop
  // body serves all of them (a single copy of the score code matters: these kernels are bound by
  // instruction fetch).  An arg-max on a border column / row is recorded as xb / xe (yb / ye).
  const int xb = (int)(x_1 + 1), xe = (int)x1, yb = (int)(y_1 + 1), ye = (int)y1;
  const int ncols = (xe >= xb ? xe - xb + 1 : 0) + 2, nrows = (ye >= yb ? ye - yb + 1 : 0) + 2;
  int max_x = xb, max_y = yb, n = 0;
  float max = 0.0f;
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
  for (int r = 0; r < nrows; ++r) {
    const bool yi = r > 0 && r < nrows - 1;
    const float yf = r == 0 ? y_1 : (yi ? (float)(yb + r - 1) : y1);
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
    for (int c = 0; c < ncols; ++c) {
      const bool xi = c > 0 && c < ncols - 1;
      const float xf = c == 0 ? x_1 : (xi ? (float)(xb + c - 1) : x1);
      const int px = xb + c - 1, py = yb + r - 1;  // the pixel itself on interior positions
      const float tmp = (xi && yi) ? (float)nb.get(px, py) : (float)tile_score_f(nb, xf, yf);
      ++n;
      if (r < nrows - 1 && tmp > (float)threshold) { *steps = n; return false; }  // not tested on the bottom row
      if (BELOW && xi && yi && tmp == max) {
        // tie rule of GetScoreMaxBelow (:987-1010): the larger 3x3 binomial sum (centre excluded) wins
        const int t1 = 2 * (nb.get(px - 1, py) + nb.get(px + 1, py) + nb.get(px, py + 1) + nb.get(px, py - 1)) +
                       (nb.get(px + 1, py + 1) + nb.get(px - 1, py + 1) + nb.get(px + 1, py - 1) + nb.get(px - 1, py - 1));
        const int t2 = 2 * (nb.get(max_x - 1, max_y) + nb.get(max_x + 1, max_y) + nb.get(max_x, max_y + 1) + nb.get(max_x, max_y - 1)) +
                       (nb.get(max_x + 1, max_y + 1) + nb.get(max_x - 1, max_y + 1) + nb.get(max_x + 1, max_y - 1) + nb.get(max_x - 1, max_y - 1));
        if (t1 > t2) { max_x = px; max_y = py; }
      }
      if (n == 1 || tmp > max) {
        max = tmp;
        if (n > 1) {
          max_x = c == 0 ? xb : (xi ? px : xe);
          if (r > 0) max_y = yi ? py : ye;
        }
      }
    }
  }
  *max_out = max; *mx = max_x; *my = max_y; *steps = n;
  return true;
}

BRISK_HD bool saturate1(float* dx, float* dy) {
  bool inside = true;
  if (*dx > 1.0f) { *dx = 1.0f; inside = false; }
  if (*dx < -1.0f) { *dx = -1.0f; inside = false; }
  if (*dy > 1.0f) { *dy = 1.0f; inside = false; }
  if (*dy < -1.0f) { *dy = -1.0f; inside = false; }
  return inside;
}

// Patch of the layer above that GetScoreMaxAbove scans for a corner of `layer`.
BRISK_HD void above_patch(int layer, int x, int y, float* x_1, float* x1, float* y_1, float* y1) {
  if ((layer & 1) == 0) {
    *x_1 = (float)((double)(float)(4 * x - 1 - 2) / 6.0); *x1 = (float)((double)(float)(4 * x - 1 + 2) / 6.0);
    *y_1 = (float)((double)(float)(4 * y - 1 - 2) / 6.0); *y1 = (float)((double)(float)(4 * y - 1 + 2) / 6.0);
  } else {
    *x_1 = (float)(6 * x - 1 - 3) / 8.0f; *x1 = (float)(6 * x - 1 + 3) / 8.0f;
    *y_1 = (float)(6 * y - 1 - 3) / 8.0f; *y1 = (float)(6 * y - 1 + 3) / 8.0f;
  }
}

// Cache footprint of a GetScoreMaxAbove call: how many patch positions it
// evaluated and, when the scan ran to completion, the arg-max whose 3x3
// neighbourhood is looked up afterwards.
struct AboveFootprint {
  int steps;      // evaluated patch positions (>= 1)
  int completed;  // scan not aborted by the threshold test
  int mx, my;     // arg-max (valid when completed)
};

// Patch of the layer below that GetScoreMaxBelow scans for a corner of `layer` (:917-933).
BRISK_HD void below_patch(int layer, int x, int y, float* x_1, float* x1, float* y_1, float* y1) {
  if ((layer & 1) == 0) {
    *x_1 = (float)((double)(float)(8 * x + 1 - 4) / 6.0); *x1 = (float)((double)(float)(8 * x + 1 + 4) / 6.0);
    *y_1 = (float)((double)(float)(8 * y + 1 - 4) / 6.0); *y1 = (float)((double)(float)(8 * y + 1 + 4) / 6.0);
  } else {
    *x_1 = (float)((double)(float)(6 * x + 1 - 3) / 4.0); *x1 = (float)((double)(float)(6 * x + 1 + 3) / 4.0);
    *y_1 = (float)((double)(float)(6 * y + 1 - 3) / 4.0); *y1 = (float)((double)(float)(6 * y + 1 + 3) / 4.0);
  }
}

// GetScoreMaxAbove (brisk-scale-space.cc:757-915) / GetScoreMaxBelow (:917-1099) in one body (the two
// differ in the patch, the tie rule of the scan and the mapping back; one body keeps the kernels' code
// small enough for the instruction cache).  `layer` is the index of the corner's own layer, `nb` the
// neighbouring layer.  Pure; the cache footprint of the scan is returned (it matters for the layer above
// only: look-ups on the layer below land on a layer whose own NMS is already finished).
BRISK_HD float score_max_side(bool below, const LayerView& nb, int layer, int x, int y, int thr, bool* ismax, float* dx,
                              float* dy, AboveFootprint* fp, int* tile_violations, const unsigned long long* pre = nullptr) {
  *ismax = false;
  float x_1, x1, y_1, y1;
  if (below) below_patch(layer, x, y, &x_1, &x1, &y_1, &y1);
  else above_patch(layer, x, y, &x_1, &x1, &y_1, &y1);
  ScoreTile tile;
  if (pre) load_scan_tile(pre, x_1, x1, y_1, y1, &tile);
  else fill_scan_tile(nb, x_1, x1, y_1, y1, &tile);
  float max; int mx = 0, my = 0, steps = 0;
  fp->completed = 0; fp->mx = 0; fp->my = 0;
  const bool ok = scan_patch(below, tile, x_1, x1, y_1, y1, thr + kDropThreshold, &max, &mx, &my, &steps);
  fp->steps = steps;
  float refined = 0.0f, dx1 = 0.0f, dy1 = 0.0f;
  if (ok) refined = tile_patch3x3(tile, mx, my, &dx1, &dy1);
#ifdef BRISK_TILE_CHECK
  if (tile_violations) *tile_violations += tile.violations;
#else
  (void)tile_violations;
#endif
  if (!ok) return 0.0f;
  fp->completed = 1; fp->mx = mx; fp->my = my;
  const float rx = (float)mx + dx1, ry = (float)my + dy1;
  const bool octave = (layer & 1) == 0;
  if (!below) {
    if (octave) {
      *dx = (rx * 6.0f + 1.0f) / 4.0f - (float)x;
      *dy = (ry * 6.0f + 1.0f) / 4.0f - (float)y;
    } else {
      *dx = (float)(((double)rx * 8.0 + 1.0) / 6.0 - (double)(float)x);
      *dy = (float)(((double)ry * 8.0 + 1.0) / 6.0 - (double)(float)y);
    }
  } else if (octave) {
    *dx = (float)(((double)rx * 6.0 + 1.0) / 8.0 - (double)(float)x);
    *dy = (float)(((double)ry * 6.0 + 1.0) / 8.0 - (double)(float)y);
  } else {
    *dx = (float)(((double)rx * 4.0 - 1.0) / 6.0 - (double)(float)x);
    *dy = (float)(((double)ry * 4.0 - 1.0) / 6.0 - (double)(float)y);
  }
  const bool inside = saturate1(dx, dy);
  *ismax = true;
  return inside ? (refined > max ? refined : max) : max;
}

// Marks, in the touch map of the layer above, the pixels that the first
// `steps` positions of scan_patch's fixed visiting order look up (float
// positions read the 2x2 cell they fall in), plus the 3x3 patch around the
// arg-max when the scan completed.  No scores are evaluated.
BRISK_HD void mark_px(const LayerView& nb, int x, int y) {
  if (!in_border(nb, x, y)) nb.bm[(long long)y * nb.pitch + x] = 1;
}
BRISK_HD void mark_cell(const LayerView& nb, float xf, float yf) {
  const int x = (int)xf, y = (int)yf;
  mark_px(nb, x, y); mark_px(nb, x + 1, y); mark_px(nb, x, y + 1); mark_px(nb, x + 1, y + 1);
}
BRISK_HD void replay_scan_marks(const LayerView& nb, float x_1, float x1, float y_1, float y1, const AboveFootprint& fp) {
  const int xb = (int)(x_1 + 1), xe = (int)x1, yb = (int)(y_1 + 1), ye = (int)y1;
  int n = fp.steps;
#define BRISK_STEP(stmt) { stmt; if (--n == 0) goto done; }
  BRISK_STEP(mark_cell(nb, x_1, y_1))
  for (int x = xb; x <= xe; ++x) BRISK_STEP(mark_cell(nb, (float)x, y_1))
  BRISK_STEP(mark_cell(nb, x1, y_1))
  for (int y = yb; y <= ye; ++y) {
    BRISK_STEP(mark_cell(nb, x_1, (float)y))
    for (int x = xb; x <= xe; ++x) BRISK_STEP(mark_px(nb, x, y))
    BRISK_STEP(mark_cell(nb, x1, (float)y))
  }
  BRISK_STEP(mark_cell(nb, x_1, y1))
  for (int x = xb; x <= xe; ++x) BRISK_STEP(mark_cell(nb, (float)x, y1))
  BRISK_STEP(mark_cell(nb, x1, y1))
#undef BRISK_STEP
done:
  if (fp.completed)
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) mark_px(nb, fp.mx + dx, fp.my + dy);
}

// ---------------------------------------------------------------------------
// Phase 1: the eight comparisons of IsMax2D (brisk-scale-space.cc:437-462).
// Order independent.  Updates the corner's map entry; for tying corners also
// stores the FAST scores of the 5x5 neighbourhood (row-major, clipped at 0;
// entries of border pixels and of corners are unused) for phase 3.
// ---------------------------------------------------------------------------
// One entry of the 5x5 score window around a corner (row-major, offsets -2..2): on the eight neighbours the value
// IsMax2D's look-up returns (T at a corner, else the FAST score clipped at 0), on the outer ring the FAST score
// (0 at corners: their cache byte is their T, which tie_pixel_value takes from the corner map), 0 at the centre
// and on border pixels.  `t` = corner-map T of the pixel, F = its clipped FAST score (only read where t == 0).
BRISK_HD int window_value(int wx, int wy, bool border, int t, int F) {
  if ((wx == 0 && wy == 0) || border) return 0;
  if (wx >= -1 && wx <= 1 && wy >= -1 && wy <= 1) return t ? t : F;
  return t ? 0 : F;
}

BRISK_HD void own_window(const LayerView& L, int x, int y, uint8_t own[25]) {
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
  for (int i = 0; i < 25; ++i) {
    const int wx = i % 5 - 2, wy = i / 5 - 2;
    const int qx = x + wx, qy = y + wy;
    const bool border = in_border(L, qx, qy);
    const int t = border ? 0 : (L.cm[(long long)qy * L.pitch + qx] & kCmT);
    const int F = (border || t || (wx == 0 && wy == 0)) ? 0 : fastF(L, qx, qy);
    own[i] = (uint8_t)window_value(wx, wy, border, t, F);
  }
}

// The eight comparisons from the window: corner-map entry of the corner (calls = look-ups made up to and including
// the first neighbour that exceeds the centre).
BRISK_HD uint16_t prefix_entry(int center, const uint8_t own[25]) {
  int calls = 8;
  bool tie = false, rejected = false;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int j = 0; j < 8; ++j) {
    int dx, dy;
    isMax2dOffset(j, &dx, &dy);
    const int v = own[(dy + 2) * 5 + dx + 2];
    if (!rejected) {
      if (v > center) { calls = j + 1; rejected = true; }
      else if (v == center) tie = true;
    }
  }
  uint16_t e = (uint16_t)(center | (calls << kCmCallsShift));
  if (rejected) e |= kCmDecided;
  else if (!tie) e |= (uint16_t)(kCmDecided | kCmAccept);
  else e |= kCmTie;
  return e;
}

#ifdef __CUDACC__
// Rows [r0, r1] of the window (offsets -2..2) from packed row evaluations: six pixels x-2..x+3 per row (one copy of the
// evaluator's code for all rows; a row's five bytes go to its own 64-bit register, so nothing is indexed dynamically).
struct OwnWindowRegs { unsigned long long r[5]; };
__device__ __forceinline__ void own_window_rows(const LayerView& L, int x, int y, int r0, int r1, OwnWindowRegs* w) {
#pragma unroll 1
  for (int wy = r0; wy <= r1; ++wy) {
    const int qy = y + wy;
    uint32_t f[3];
    fast916_row<3>(L.img, L.pitch, L.h, x - 2, qy, f);
    unsigned long long rowv = 0;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const int wx = j - 2, qx = x + wx;
      const bool border = in_border(L, qx, qy);
      const int t = border ? 0 : (L.cm[(long long)qy * L.pitch + qx] & kCmT);
      // (wy is only used for "inner 3x3 or not": rows -1..1 all behave like row 0 except for the centre)
      const int v = window_value(wx, wy == 0 ? 0 : (wy == -1 || wy == 1 ? 1 : 2), border, t, (int)((f[j >> 1] >> ((j & 1) * 16)) & 0xffu));
      rowv |= (unsigned long long)v << (8 * j);
    }
    if (wy == -2) w->r[0] = rowv;
    else if (wy == -1) w->r[1] = rowv;
    else if (wy == 0) w->r[2] = rowv;
    else if (wy == 1) w->r[3] = rowv;
    else w->r[4] = rowv;
  }
}
#endif

// Window + corner-map entry of one corner.  Host: the whole window.  Device (nms_prefix_kernel): the three middle rows
// first (they hold the eight neighbours) with nms_prefix_mid; the outer rows only for corners that tie -- nothing else
// reads them -- and those are evaluated by the warp together, one row per lane (own_window_rows on the owner's data).
BRISK_HD void nms_prefix(const LayerView& L, int x, int y, uint8_t fwin[25]) {
  const long long o = (long long)y * L.pitch + x;
  own_window(L, x, y, fwin);
  L.cm[o] = prefix_entry(L.cm[o] & kCmT, fwin);
}
#ifdef __CUDACC__
__device__ __forceinline__ uint16_t nms_prefix_mid(const LayerView& L, int x, int y, OwnWindowRegs* w) {
  const long long o = (long long)y * L.pitch + x;
  w->r[0] = w->r[1] = w->r[2] = w->r[3] = w->r[4] = 0;
  own_window_rows(L, x, y, -1, 1, w);
  uint8_t mid[25];
#pragma unroll
  for (int i = 0; i < 25; ++i) mid[i] = (i >= 5 && i < 20) ? (uint8_t)(w->r[i / 5] >> (8 * (i % 5))) : (uint8_t)0;
  const uint16_t e = prefix_entry(L.cm[o] & kCmT, mid);
  L.cm[o] = e;
  return e;
}
#endif

// Scores of the two scan tiles of a corner (the layer above: bytes 0..15, the layer below: bytes 16..31, row stride 4
// from the tile's first pixel; for layer 0 bytes 16..24 hold the nine 5-8 scores of :558-592, s[3 * column + row]).
// Scalar statement of what nms_windows_kernel computes with one lane per byte.
BRISK_HD int side_tile_value(const LayerView& below, const LayerView& L, const LayerView& above, int n_layers, int layer, int x, int y,
                             int i) {
  const bool lower = i >= 16;
  const int l = i & 15;
  if (lower && layer == 0) return l < 9 ? score58(L, x + l / 3 - 1, y + l % 3 - 1) : 0;
  if (lower ? layer == 0 : layer == n_layers - 1) return 0;
  float x_1, x1, y_1, y1;
  if (lower) below_patch(layer, x, y, &x_1, &x1, &y_1, &y1);
  else above_patch(layer, x, y, &x_1, &x1, &y_1, &y1);
  int xa, ya, xb, yb;
  scan_tile_bounds(x_1, x1, y_1, y1, &xa, &ya, &xb, &yb);
  const int tx = xa + (l & 3), ty = ya + (l >> 2);
  if (tx > xb || ty > yb) return 0;
  return score1(lower ? below : above, tx, ty);
}
BRISK_HD void side_tiles(const LayerView& below, const LayerView& L, const LayerView& above, int n_layers, int layer, int x, int y,
                         uint8_t tiles[32]) {
  for (int i = 0; i < 32; ++i) tiles[i] = n_layers == 1 ? 0 : (uint8_t)side_tile_value(below, L, above, n_layers, layer, x, y, i);
}

// ---------------------------------------------------------------------------
// Phase 2: scale-space checks of a corner that passed phase 1 (tying or not),
// without side effects.  Results are kept for refine_emit.
// ---------------------------------------------------------------------------
struct CheckResult {
  float max_above, dxa, dya;
  float max_below, dxb, dyb;
  int above_steps;  // AboveFootprint::steps | completed << 8 (mid layers)
  int above_argmax; // mx | my << 16
};

// Returns true when Refine3D (mid layers) / the last-layer branch of
// GetKeypoints reaches its own-layer 3x3 patch.
// `below` / `above` are the neighbouring layers (ignored where the corner's layer has none).
// `center_in` < 0: the corner's own score is its corner-map entry (detection mode); otherwise the caller
// passes GetAgastScore(x, y, 1) of a point that need not be a corner (provided-key-point mode).
BRISK_HD bool nms_checks3(const LayerView& below, const LayerView& L, const LayerView& above, int n_layers, int layer, int x,
                          int y, CheckResult* r, int* tile_violations = nullptr, int center_in = -1,
                          const unsigned long long* pre = nullptr /* side_tiles as four 64-bit words */) {
  const int center = center_in >= 0 ? center_in : (L.cm[(long long)y * L.pitch + x] & kCmT);
  r->max_above = 0; r->dxa = 0; r->dya = 0; r->max_below = 0; r->dxb = 0; r->dyb = 0;
  r->above_steps = 0; r->above_argmax = 0;
  if (n_layers == 1) return true;
  // the layer above first (Refine3D :540-556; none for the last layer :213-222), then the layer below; both run
  // through one copy of the scan code
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
  for (int side = layer == n_layers - 1 ? 1 : 0; side < 2; ++side) {
    if (side == 1 && layer == 0) break;
    bool ismax;
    float dx = 0.0f, dy = 0.0f;
    AboveFootprint fp;
    const float m = score_max_side(side == 1, side == 1 ? below : above, layer, x, y, center, &ismax, &dx, &dy, &fp, tile_violations,
                                   pre ? pre + 2 * side : nullptr);
    if (side == 0) {
      r->max_above = m; r->dxa = dx; r->dya = dy;
      r->above_steps = fp.steps | (fp.completed << 8);
      r->above_argmax = fp.mx | (fp.my << 16);
      if (!ismax) return false;
    } else {
      r->max_below = m; r->dxb = dx; r->dyb = dy;
      return ismax;
    }
  }
  // layer 0: guess the virtual intra-octave below octave 0 with the 5-8 mask (:558-592)
  int s[9], best = 0;
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
  for (int k = 0; k < 9; ++k) {  // s[3 * column + row], the reference's s_<column>_<row>
    s[k] = pre ? (int)((pre[2 + (k >> 3)] >> ((k & 7) << 3)) & 0xffull) : score58(L, x + k / 3 - 1, y + k % 3 - 1);
    best = imax(best, s[k]);
  }
  subpixel2d(s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7], s[8], &r->dxb, &r->dyb);
  r->max_below = (float)best;
  return true;
}

BRISK_HD bool nms_checks(const LayerView* layers, int n_layers, int layer, int x, int y, CheckResult* r,
                         int* tile_violations = nullptr, const unsigned long long* pre = nullptr) {
  return nms_checks3(layers[layer > 0 ? layer - 1 : 0], layers[layer], layers[layer + 1 < n_layers ? layer + 1 : layer], n_layers,
                     layer, x, y, r, tile_violations, -1, pre);
}

// ---------------------------------------------------------------------------
// Phase 3: the tie path of IsMax2D (brisk-scale-space.cc:464-530) for one
// corner, given that every raster-earlier corner of the layer is decided.
// ---------------------------------------------------------------------------

// 8x8 window of corner-map entries around a tying corner (cx-4..cx+3, cy-4..cy+3):
// every raster-earlier corner that can influence the 5x5 pixels IsMax2D's tie path
// reads lies inside it.  Entries are staged once (64 independent loads) into
// `w[i * stride]` -- per-thread scratch in shared memory on the GPU -- so that the
// state reconstruction below runs on chip.
struct TieWindow {
  const uint16_t* w;
  int stride, x0, y0;
  BRISK_HD int at(int px, int py) const { return w[(((py - y0) << 3) + (px - x0)) * stride]; }
};

// Returns true when a raster-earlier corner inside the window is still
// undecided (the tying corner cannot be resolved yet).
BRISK_HD bool load_tie_window(const LayerView& L, int cx, int cy, uint16_t* w, int stride) {
  bool blocked = false;
  for (int j = 0; j < 8; ++j) {
    const int py = cy - 4 + j;
    for (int i = 0; i < 8; ++i) {
      const int px = cx - 4 + i;
      uint16_t e = 0;
      if (py >= 3 && px >= 3 && px < L.w - 3 && py < L.h - 3) e = L.cm[(long long)py * L.pitch + px];
      w[((j << 3) + i) * stride] = e;
      if ((e & kCmT) && !(e & kCmDecided) && (py < cy || (py == cy && px < cx))) blocked = true;
    }
  }
  return blocked;
}

// Raw cache byte the reference would hold at pixel (qx,qy) just before corner
// (cx,cy) runs IsMax2D.  F is the pixel's FAST score clipped at 0.  Raster-earlier
// tying corners that are not decided yet count as rejected, or as accepted when
// `assume_accept` is set; the result is monotone in those bits (an accepted corner
// can only make a pixel sticky), so a value that is the same under both assumptions
// does not depend on the pending decisions at all.
BRISK_HD int cache_state(const LayerView& L, const TieWindow& W, int mode, int qx, int qy, int F, int cx, int cy,
                         bool assume_accept = false) {
  if (in_border(L, qx, qy)) return 0;
  const int tq = W.at(qx, qy) & kCmT;
  if (tq) return tq;  // a detected corner holds its threshold-map value (> 2)
  if (F < 1) return 0;
  bool sticky = false;
  int last = 0;  // threshold of the most recent look-up, 0 = none
  if (L.bm[(long long)qy * L.pitch + qx]) { sticky = true; last = 1; }
  // look-ups by raster-earlier corners of this layer whose footprint holds q
  for (int py = qy - 2; py <= qy + 1; ++py) {
    if (py > cy) continue;
    for (int px = qx - 2; px <= qx + 1; ++px) {
      if (py == cy && px >= cx) continue;  // not earlier than (cx, cy)
      const int e = W.at(px, py);
      if (!(e & kCmT)) continue;
      const int ox = qx - px, oy = qy - py;  // in [-1,2]^2
      if (ox <= 1 && oy <= 1) {
        // IsMax2D neighbour look-up with threshold T(p), if p got that far
        const int j = isMax2dIndex(ox, oy);
        if (j < ((e & kCmCalls) >> kCmCallsShift)) {
          const int t = e & kCmT;
          if (t <= F) sticky = true;
          last = t;
        }
      }
      if ((e & kCmAccept) || (assume_accept && !(e & kCmDecided))) {
        bool touched;
        if (mode == kModeSingle) touched = true;                                       // 4x4 float-accessor patch
        else if (mode == kModeLast) touched = (ox >= 0 && oy >= 0 && ox <= 1 && oy <= 1) || (e & kCmChecks);  // 2x2 centre read, then 4x4 patch
        else touched = (e & kCmChecks) && ox <= 1 && oy <= 1;                           // int 3x3 patch of Refine3D
        if (touched) { sticky = true; last = 1; }
      }
    }
  }
  if (F > 2) return sticky ? F : 0;
  return (last != 0 && last <= F) ? F : 0;
}

// Value IsMax2D's tie path sees at offset (ox, oy) in [-2,2]^2 from the tying corner (x, y) once the
// corner's own eight neighbour look-ups are done: its own score at the centre, the look-up results
// s(q) on the 8 neighbours, raw cache bytes on the outer ring.  F = fwin entry of that pixel.
BRISK_HD int tie_pixel_value(const LayerView& L, const TieWindow& W, int mode, int x, int y, int ox, int oy, int F, int center,
                             bool assume_accept = false) {
  if (ox == 0 && oy == 0) return center;
  const int qx = x + ox, qy = y + oy;
  if (ox >= -1 && ox <= 1 && oy >= -1 && oy <= 1) {
    if (W.at(qx, qy) & kCmT) return F;  // neighbouring corner: its T (stored in fwin by nms_prefix)
    const int st = cache_state(L, W, mode, qx, qy, F, x, y, assume_accept);
    return st > 2 ? st : (F >= center ? F : 0);
  }
  return cache_state(L, W, mode, qx, qy, F, x, y, assume_accept);
}

// IsMax2D verdict of the tying corner (x,y): 1 accept, 0 reject, -1 not yet
// decidable: one of the 25 values the tie path reads depends on whether a
// raster-earlier tying corner of its neighbourhood, still undecided, will be
// accepted.  (Most pending neighbours do not matter: their patch misses the
// pixels read here, or those pixels are sticky / zero anyway.)  Decidable corners
// can be resolved in any order, or concurrently: a decision is published with one
// 16-bit store.  `scratch` holds 64 entries at `stride`.
BRISK_HD int nms_tie_decide(const LayerView& L, int mode, int x, int y, const uint8_t fwin[25], uint16_t* scratch, int stride) {
  const bool pending = load_tie_window(L, x, y, scratch, stride);
  const TieWindow W{scratch, stride, x - 4, y - 4};
  const int center = W.at(x, y) & kCmT;
  // every value is known up to the pending verdicts: v0 with all of them taken as reject, v1 as accept
  // (v0 <= v1, and the true value is one of the two)
  int v0[25], v1[25];
  for (int i = 0; i < 25; ++i) {
    const int ox = i % 5 - 2, oy = i / 5 - 2;
    v0[i] = tie_pixel_value(L, W, mode, x, y, ox, oy, fwin[i], center);
    v1[i] = pending ? tie_pixel_value(L, W, mode, x, y, ox, oy, fwin[i], center, true) : v0[i];
  }
  // 3x3 binomial sums (weights 4 / 2 / 1) around the centre and around every tying neighbour, as intervals.
  // Reject if a certainly tying neighbour certainly beats the centre, accept if no possibly tying one possibly
  // does; otherwise the verdict depends on a pending one.
  int s_lo = 0, s_hi = 0;
  for (int wy = -1; wy <= 1; ++wy)
    for (int wx = -1; wx <= 1; ++wx) {
      const int wgt = (wx == 0 ? 2 : 1) * (wy == 0 ? 2 : 1), i = (2 + wy) * 5 + 2 + wx;
      s_lo += wgt * v0[i]; s_hi += wgt * v1[i];
    }
  bool may_lose = false;
  for (int ty = -1; ty <= 1; ++ty)
    for (int tx = -1; tx <= 1; ++tx) {
      const int t = (2 + ty) * 5 + 2 + tx;
      if ((tx == 0 && ty == 0) || (v0[t] != center && v1[t] != center)) continue;
      int o_lo = 0, o_hi = 0;
      for (int wy = -1; wy <= 1; ++wy)
        for (int wx = -1; wx <= 1; ++wx) {
          const int wgt = (wx == 0 ? 2 : 1) * (wy == 0 ? 2 : 1), i = (2 + ty + wy) * 5 + 2 + tx + wx;
          o_lo += wgt * v0[i]; o_hi += wgt * v1[i];
        }
      if (v0[t] == center && v1[t] == center && o_lo > s_hi) return 0;
      if (o_hi > s_lo) may_lose = true;
    }
  return may_lose ? -1 : 1;
}

// ---------------------------------------------------------------------------
// Phase 4: cache footprint an accepted corner of `layer` leaves on the layer
// above (its GetScoreMaxAbove look-ups), recorded in that layer's touch map.
// ---------------------------------------------------------------------------
BRISK_HD void mark_above1(const LayerView& above, int layer, int x, int y, const CheckResult& r) {
  float x_1, x1, y_1, y1;
  above_patch(layer, x, y, &x_1, &x1, &y_1, &y1);
  AboveFootprint fp;
  fp.steps = r.above_steps & 0xff; fp.completed = (r.above_steps >> 8) & 1;
  fp.mx = r.above_argmax & 0xffff; fp.my = r.above_argmax >> 16;
  replay_scan_marks(above, x_1, x1, y_1, y1, fp);
}
BRISK_HD void mark_above(const LayerView* layers, int layer, int x, int y, const CheckResult& r) {
  mark_above1(layers[layer + 1], layer, x, y, r);
}

// ---------------------------------------------------------------------------
// Phase 5: refinement of an accepted corner (Refine3D :598-754 after the
// checks, and the single / last-layer branches of GetKeypoints :172-256).
// Returns false when the corner is discarded on the scale axis.
// ---------------------------------------------------------------------------
// 3x3 patch of the corner's own layer from its score window (own_window): the neighbours' look-up values and T at the centre.
BRISK_HD float patch3x3_window(const uint8_t own[25], int center, float* dx, float* dy) {
  int s[9];  // s[3 * column + row]
  for (int k = 0; k < 9; ++k) s[k] = own[(k % 3 + 1) * 5 + k / 3 + 1];
  s[4] = center;
  return subpixel2d(s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7], s[8], dx, dy);
}

BRISK_HD bool refine_emit1(const LayerView& L, int n_layers, int layer, int x, int y, const CheckResult& r, KeyPoint* kp,
                           const uint8_t* own = nullptr) {
  float dxl, dyl;
  int s11;
  float max_layer;
  if (own) { s11 = L.cm[(long long)y * L.pitch + x] & kCmT; max_layer = patch3x3_window(own, s11, &dxl, &dyl); }
  else max_layer = patch3x3(L, x, y, &dxl, &dyl, &s11);
  kp->angle = -1.0f; kp->class_id = -1; kp->octave = layer;
  if (n_layers == 1) {
    kp->x = (float)x + dxl; kp->y = (float)y + dyl; kp->size = 12.0f; kp->response = max_layer; kp->octave = 0;
    return true;
  }
  if (layer == n_layers - 1) {
    kp->x = ((float)x + dxl) * L.scale + L.offset; kp->y = ((float)y + dyl) * L.scale + L.offset;
    kp->size = 12.0f * L.scale; kp->response = max_layer;
    return true;
  }
  const int center = s11;  // == T at the corner
  const bool octave = (layer & 1) == 0;
  bool refine_scale = true;
  if (layer == 0) {
    if (s11 - kMaxThreshold <= (int)r.max_above) refine_scale = false;
  } else if ((float)(s11 - kMaxThreshold) < r.max_above || (float)(s11 - kMaxThreshold) < r.max_below) {
    if ((float)(s11 - kMinDrop) > r.max_above || (float)(s11 - kMinDrop) > r.max_below) refine_scale = false;
    else return false;
  }
  float scale, max;
  if (refine_scale) {
    const int kind = octave ? (layer == 0 ? 2 : 0) : 1;
    const float c = (float)center;
    scale = refine1d(kind, r.max_below, c > max_layer ? c : max_layer, r.max_above, &max);
  } else {
    scale = 1.0f;
    max = max_layer;
  }
  float r0, ox, oy;
  if (octave) {
    if (scale > 1.0f) { r0 = (float)((1.5 - (double)scale) / .5); ox = r.dxa; oy = r.dya; }
    else if (layer == 0) { r0 = (float)(((double)scale - 0.5) / 0.5); ox = r.dxb; oy = r.dyb; }
    else { r0 = (float)(((double)scale - 0.75) / 0.25); ox = r.dxb; oy = r.dyb; }
  } else {
    if (scale > 1.0f) { r0 = (float)(4.0 - (double)scale * 3.0); ox = r.dxa; oy = r.dya; }
    else { r0 = (float)((double)scale * 3.0 - 2.0); ox = r.dxb; oy = r.dyb; }
  }
  const float r1 = (float)(1.0 - (double)r0);
  const float px = (r0 * dxl + r1 * ox) + (float)x;
  const float py = (r0 * dyl + r1 * oy) + (float)y;
  if (layer == 0 && !(scale > 1.0f)) { kp->x = px; kp->y = py; }
  else { kp->x = px * L.scale + L.offset; kp->y = py * L.scale + L.offset; }
  kp->size = 12.0f * (scale * L.scale);
  kp->response = max;
  return true;
}
BRISK_HD bool refine_emit(const LayerView* layers, int n_layers, int layer, int x, int y, const CheckResult& r,
                          KeyPoint* kp, const uint8_t* own = nullptr) {
  return refine_emit1(layers[layer], n_layers, layer, x, y, r, kp, own);
}

// ---------------------------------------------------------------------------
// "Provided key points" mode: BriskFeatureDetector::ComputeScale (reference
// brisk-feature-detector.cc:87-92), i.e. BriskScaleSpace::GetKeypoints with a
// non-empty key-point vector (brisk-scale-space.cc:104-124) on a pyramid whose
// lower threshold is 0.  No detector runs and IsMax2D is skipped; what is left
// of the lazy score cache is again a pure function once three things are known
// per pixel q of a layer (all of it happens before the first refinement):
//   * q is one of the four bilinear neighbours of a provided point: it was
//     scored with threshold 0, and a pixel that is no corner at b = 0 then holds
//     cornerScore = -1 stored as the byte 255 -- sticky (brisk-layer.cc:124-130);
//   * q is the flat offset int(x + y * cols) of a provided point (float
//     arithmetic, brisk-layer.cc:110-116): it holds the threshold-map value T,
//     sticky when T > 2 (and T >= F, so otherwise nothing sticks there);
//   * else the usual F >= 1 ? F : 0.
// The corner map carries these as its T byte (255 / T / 0), so score1() serves
// every look-up unchanged.
// ---------------------------------------------------------------------------

// Layer coordinates of a provided key point and the border test of :110-121.
BRISK_HD bool provided_to_layer(const LayerView& L, float kx, float ky, float* px, float* py) {
  const float x = kx / L.scale - L.offset, y = ky / L.scale - L.offset;
  *px = x; *py = y;
  return !(x < 3 || y < 3 || x > (float)(L.w - 3) || y > (float)(L.h - 3));
}

// Pass 1 (any order, idempotent): the threshold-0 look-ups of GetAgastScore(x, y, 0) (:119).
BRISK_HD void provided_touch(const LayerView& L, float px, float py) {
  const int x = (int)px, y = (int)py;
  for (int k = 0; k < 4; ++k) {
    const int qx = x + (k & 1), qy = y + (k >> 1);
    if (in_border(L, qx, qy)) continue;
    if (fast916(L.img, L.pitch, qx, qy) < 0) L.cm[(long long)qy * L.pitch + qx] = 255;
  }
}

// Pass 2 (after pass 1 of every point of the layer; idempotent): the score write of GetAgastPoints.
BRISK_HD void provided_stamp(const LayerView& L, float px, float py) {
  const int offs = (int)(px + py * (float)L.w);
  const int qx = offs % L.w, qy = offs / L.w;
  if (in_border(L, qx, qy)) return;  // never read back (GetAgastScore's border test comes first)
  const int T = thrmap_px(L.img, L.pitch, qx, qy);
  L.cm[(long long)qy * L.pitch + qx] = (uint16_t)(T > 2 ? T : 0);
}

// BriskLayer::GetAgastScore(float, float, 1) without a tile.
BRISK_HD int score1_f(const LayerView& L, float xf, float yf) {
  const int x = (int)xf;
  const float rx1 = xf - (float)x;
  const float rx = 1.0f - rx1;
  const int y = (int)yf;
  const float ry1 = yf - (float)y;
  const float ry = 1.0f - ry1;
  const int s00 = score1(L, x, y), s10 = score1(L, x + 1, y), s01 = score1(L, x, y + 1), s11 = score1(L, x + 1, y + 1);
  const float v = ((((rx * ry) * (float)s00 + (rx1 * ry) * (float)s10) + (rx * ry1) * (float)s01) + (rx1 * ry1) * (float)s11);
  return (int)(uint8_t)v;
}

// The 3x3 patch of :184-195 / :232-243 read at float positions.
BRISK_HD float patch3x3_f(const LayerView& L, float x, float y, float* dx, float* dy) {
  int s[9];  // s[3 * column + row]
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
  for (int k = 0; k < 9; ++k) s[k] = score1_f(L, x + (float)(k / 3 - 1), y + (float)(k % 3 - 1));
  return subpixel2d(s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7], s[8], dx, dy);
}

// Pass 3 (pure): one provided point of one layer -> at most one key point (:172-285 with
// perform_2d_nonMax = false).  class_id is the caller's (the reference copies the input key point).
BRISK_HD bool provided_refine(const LayerView& below, const LayerView& L, const LayerView& above, int n_layers, int layer,
                              float px, float py, KeyPoint* kp, int* tile_violations = nullptr) {
  kp->angle = -1.0f; kp->octave = layer;
  if (n_layers == 1) {
    float dx, dy;
    const float m = patch3x3_f(L, px, py, &dx, &dy);
    kp->x = px + dx; kp->y = py + dy; kp->size = 12.0f; kp->response = m; kp->octave = 0;
    return true;
  }
  const int x = (int)px, y = (int)py;
  if (layer == n_layers - 1) {
    bool ismax;
    float dx = 0.0f, dy = 0.0f;
    AboveFootprint fp;
    score_max_side(true, below, layer, x, y, score1_f(L, px, py), &ismax, &dx, &dy, &fp, tile_violations);
    if (!ismax) return false;
    const float m = patch3x3_f(L, px, py, &dx, &dy);
    kp->x = (px + dx) * L.scale + L.offset; kp->y = (py + dy) * L.scale + L.offset;
    kp->size = 12.0f * L.scale; kp->response = m;
    return true;
  }
  CheckResult r;
  if (!nms_checks3(below, L, above, n_layers, layer, x, y, &r, tile_violations, score1(L, x, y))) return false;
  const int class_id = kp->class_id;
  const bool ok = refine_emit1(L, n_layers, layer, x, y, r, kp);
  kp->class_id = class_id;
  return ok;
}

}  // namespace briskb200
