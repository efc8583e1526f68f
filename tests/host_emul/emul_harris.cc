// TEST HARNESS: the product's Harris scale-space logic (harris_logic.cuh, incl. the libstdc++
// introsort replay) compiled for the host and run serially, to be compared with the oracle.
#include <stdint.h>
#include <string.h>
#include <vector>

#include "../../ethzasl_brisk_b200/csrc/brisk_math.cuh"
#include "../../ethzasl_brisk_b200/csrc/harris_logic.cuh"

using namespace briskb200;

namespace {
struct HL {
  int w, h;
  bool octave;
  double scale, offset, scale_above, offset_above, scale_below, offset_below;
  std::vector<uint8_t> img;
  std::vector<int> sc;
};
}  // namespace

namespace {
// harris_sort_kernel (gcc_sort.cuh: libstdc++'s std::sort replayed) followed by harris_uniformity_kernel (radius > 0) or
// harris_bucketing_kernel (radius <= 0; run here point by point, the kernel takes the sorted points 32 at a time with the
// same per-bucket quota): the thinning step of one layer.
std::vector<HPoint> thin_points(std::vector<HPoint> pts, double radius, int w, int h, long long max_kpt) {
  std::vector<HPoint> keep;
  if (pts.empty()) return keep;
  gcc_sort(pts.data(), (int)pts.size());
  if (!(radius > 0.0)) {
    const unsigned step_u = 1u + (unsigned)(w - 1) / 4u, step_v = 1u + (unsigned)(h - 1) / 4u;
    const unsigned quota = (unsigned)((unsigned long long)max_kpt / 16ull);
    unsigned cnt[16] = {};
    for (const HPoint& p : pts) {
      const unsigned b = (p.x / step_u) * 4u + p.y / step_v;
      if (cnt[b] < quota) { ++cnt[b]; keep.push_back(p); }
    }
    return keep;
  }
  const float max_score = (float)pts[0].score;
  const float scaling = (float)(15.0 / (double)(float)radius);
  const int orows = (int)(h * ceil((double)scaling) + 32), ocols = (int)(w * ceil((double)scaling) + 32);
  std::vector<uint8_t> occ((size_t)orows * ocols + 64, 0);
  for (const HPoint& p : pts) {
    const int cy = (int)((float)(int)p.y * scaling + 16.0f), cx = (int)((float)(int)p.x * scaling + 16.0f);
    const float nsc1 = uniformity_nsc1(p.score, max_score);
    if ((double)nsc1 < (double)occ[(size_t)cy * ocols + cx]) continue;
    const float nsc = 0.99f * nsc1;
    for (int y = 0; y < 31; ++y)
      for (int x = 0; x < 31; ++x) {
        uint8_t& o = occ[(size_t)(cy + y - 15) * ocols + cx + x - 15];
        const int v = (int)o + uniformity_stamp(x, y, nsc);
        o = (uint8_t)(v > 255 ? 255 : v);
      }
    keep.push_back(p);
    if ((long long)keep.size() == max_kpt) break;
  }
  return keep;
}
}  // namespace

extern "C" int emul_harris_detect(const uint8_t* image, int w, int h, int octaves, double radius, double abs_thr,
                                  long long max_kpt, KeyPoint* out, int cap) {
  const int n = octaves * 2 > 1 ? octaves * 2 : 1;
  std::vector<HL> L(n);
  for (int i = 0; i < n; ++i) {
    HL& l = L[i];
    if (i == 0) { l.w = w; l.h = h; l.img.assign(image, image + (size_t)w * h); }
    else if (i == 1) {
      l.w = 2 * (w / 3); l.h = 2 * (h / 3); l.img.resize((size_t)l.w * l.h);
      for (int R = 0; R < h / 3; ++R)
        for (int T = 0; T < w / 3; ++T) {
          int p[9], o[4];
          for (int k = 0; k < 9; ++k) p[k] = L[0].img[(size_t)(3 * R + k / 3) * w + 3 * T + k % 3];
          twothird_block(p, T, w, o);
          l.img[(size_t)(2 * R) * l.w + 2 * T] = o[0]; l.img[(size_t)(2 * R) * l.w + 2 * T + 1] = o[1];
          l.img[(size_t)(2 * R + 1) * l.w + 2 * T] = o[2]; l.img[(size_t)(2 * R + 1) * l.w + 2 * T + 1] = o[3];
        }
    } else {
      const HL& s = L[i - 2];
      l.w = s.w / 2; l.h = s.h / 2; l.img.resize((size_t)l.w * l.h);
      for (int r = 0; r < l.h; ++r)
        for (int c = 0; c < l.w; ++c) {
          const uint8_t* a = &s.img[(size_t)(2 * r) * s.w + 2 * c];
          l.img[(size_t)r * l.w + c] = (uint8_t)halfsample_px(a[0], a[1], a[s.w], a[s.w + 1], c, s.w);
        }
    }
    l.octave = (i % 2 == 0);
    if (l.octave) { l.offset_above = -0.25; l.offset_below = 1.0 / 6.0; l.scale_above = 2.0 / 3.0; l.scale_below = 4.0 / 3.0; l.scale = i == 0 ? 1.0 : pow(2.0, (double)(i / 2)); }
    else { l.offset_above = -1.0 / 6.0; l.offset_below = 0.125; l.scale_above = 0.75; l.scale_below = 1.5; l.scale = pow(2.0, (double)(i / 2)) * 1.5; }
    l.offset = i == 0 ? 0.0 : l.scale * 0.5 - 0.5;
    // scores
    l.sc.assign((size_t)l.w * l.h, 0);
    std::vector<int> xx((size_t)l.w * l.h, 0), yy(xx), xy(xx);
    for (int y = 1; y < l.h - 1; ++y)
      for (int x = 1; x < l.w - 1; ++x) {
        int p[9];
        for (int k = 0; k < 9; ++k) p[k] = l.img[(size_t)(y - 1 + k / 3) * l.w + x - 1 + k % 3];
        harris_products(p, &xx[(size_t)y * l.w + x], &yy[(size_t)y * l.w + x], &xy[(size_t)y * l.w + x]);
      }
    for (int y = 2; y < l.h - 2; ++y)
      for (int x = 2; x < l.w - 2; ++x) {
        int qa[9], qb[9], qc[9];
        for (int k = 0; k < 9; ++k) {
          const size_t o = (size_t)(y - 1 + k / 3) * l.w + x - 1 + k % 3;
          qa[k] = xx[o]; qb[k] = yy[o]; qc[k] = xy[o];
        }
        l.sc[(size_t)y * l.w + x] = harris_response(harris_smooth(qa), harris_smooth(qb), harris_smooth(qc));
      }
  }
  int total = 0;
  for (int i = 0; i < n; ++i) {
    HL& l = L[i];
    std::vector<HPoint> pts;
    const int thr = (int)abs_thr;
    for (int y = 2; y < l.h - 2; ++y)
      for (int x = 2; x < l.w - 2; ++x) {
        const int* p = &l.sc[(size_t)y * l.w + x];
        const int c = *p;
        if (c < thr) continue;
        if (p[1] > c || p[-1] > c || p[l.w] > c || p[-l.w] > c || p[l.w + 1] > c || p[l.w - 1] > c || p[-l.w + 1] > c || p[-l.w - 1] > c) continue;
        pts.push_back(HPoint{c, (unsigned short)x, (unsigned short)y});
      }
    const HL* above = i + 1 < n ? &L[i + 1] : nullptr;
    const HL* below = i > 0 ? &L[i - 1] : nullptr;
    if (above || below) {
      std::vector<HPoint> kept;
      static const int off[9][2] = {{0, 0}, {1, 0}, {-1, 0}, {0, 1}, {0, -1}, {1, 1}, {1, -1}, {-1, 1}, {-1, -1}};
      for (const HPoint& p : pts) {
        bool ok = true;
        if (above)
          for (int k = 0; k < 9 && ok; ++k) {
            const double u = (int)p.x + off[k][0], v = (int)p.y + off[k][1];
            if ((double)p.score < harris_score_bilinear(above->sc.data(), above->w, above->w, above->h, l.scale_above * (u + l.offset_above), l.scale_above * (v + l.offset_above))) ok = false;
          }
        if (below && ok) {
          const double u = p.x, v = p.y;  // int(1/scale_below) == 0: nine identical reads
          if ((double)p.score < harris_score_bilinear(below->sc.data(), below->w, below->w, below->h, l.scale_below * (u + l.offset_below), l.scale_below * (v + l.offset_below))) ok = false;
        }
        if (ok) kept.push_back(p);
      }
      pts.swap(kept);
    }
    if (pts.empty()) continue;
    std::vector<HPoint> keep = thin_points(pts, radius, l.w, l.h, max_kpt);
    for (const HPoint& p : keep) {
      auto S = [&](int u, int v) { return (double)l.sc[(size_t)v * l.w + u]; };
      const int u = p.x, v = p.y;
      float dx, dy;
      harris_subpixel2d(S(u - 1, v - 1), S(u, v - 1), S(u + 1, v - 1), S(u - 1, v), S(u, v), S(u + 1, v), S(u - 1, v + 1), S(u, v + 1), S(u + 1, v + 1), &dx, &dy);
      KeyPoint k;
      k.x = (float)(l.scale * ((double)((float)(int)p.x + dx) + l.offset));
      k.y = (float)(l.scale * ((double)((float)(int)p.y + dy) + l.offset));
      k.size = (float)(l.scale * 12.0); k.angle = -1; k.response = (float)p.score; k.octave = i / 2; k.class_id = -1;
      if (total < cap) out[total] = k;
      ++total;
    }
  }
  return total;
}

// detect() on a non-empty key-point vector, one layer (harris_import_kernel, sort, thinning, harris_emit_passed_kernel).
// Returns -4 for a passed point outside the image (the C ABI reports BRISK_ERR_INVALID).
extern "C" int emul_harris_passed(int w, int h, double radius, long long max_kpt, const KeyPoint* in, int n_in, KeyPoint* out, int cap) {
  std::vector<HPoint> pts;
  for (int j = 0; j < n_in; ++j) {
    const KeyPoint& kp = in[j];
    if (!((double)kp.response > 1e6)) continue;
    if (!(kp.response < 2147483648.0f) || !(kp.x >= 0.0f && kp.x < (float)w && kp.y >= 0.0f && kp.y < (float)h)) return -4;
    pts.push_back(HPoint{(int)kp.response, (unsigned short)(int)kp.x, (unsigned short)(int)kp.y});
  }
  if (pts.empty()) {
    for (int j = 0; j < n_in && j < cap; ++j) out[j] = in[j];
    return n_in;
  }
  const std::vector<HPoint> keep = thin_points(pts, radius, w, h, max_kpt);
  int total = 0;
  for (const HPoint& p : keep) {
    KeyPoint k;
    k.x = (float)(int)p.x; k.y = (float)(int)p.y; k.size = 12.0f; k.angle = -1.0f; k.response = (float)p.score; k.octave = 0; k.class_id = -1;
    if (total < cap) out[total] = k;
    ++total;
  }
  return total;
}

