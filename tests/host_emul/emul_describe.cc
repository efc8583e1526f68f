// TEST HARNESS: the product's descriptor sampling logic (describe_logic.cuh)
// compiled for the host, run serially over key points.  Pattern tables come
// from the product's host-side builder (pattern.cc).  Not linked by the product.
#include <stdint.h>
#include <string.h>
#include <vector>

#include "../../ethzasl_brisk_b200/csrc/describe_logic.cuh"
#include "../../ethzasl_brisk_b200/csrc/pattern.h"

using namespace briskb200;

// The whole descriptor path as the GPU runs it (describe_cull_kernel + describe_kernel, describe.cu), serially:
// scale index from the tabulated size breaks, border cull (stable), integral image in the sampler's 2x2-block
// layout, orientation from the long pairs, rotated sampling, bit packing.  kps is in/out (culled, angle written);
// desc receives n_out rows of desc_bytes.  Returns n_out, or -1 when the pattern cannot be built.
extern "C" int emul_describe(const uint8_t* img, int w, int h, KeyPoint* kps, int n, int rot_inv, int scale_inv, int version,
                             float pattern_scale, uint8_t* desc, int* desc_bytes_out) {
  static PatternHost ph;
  static int have_v = -1; static float have_ps = -1;
  if (have_v != version || have_ps != pattern_scale) { if (!build_pattern(version, pattern_scale, nullptr, &ph).empty()) return -1; have_v = version; have_ps = pattern_scale; }
  const int P = ph.n_points, nb = ph.desc_bytes;
  const int n_short = (int)ph.short_pairs.size() / 2, n_long = (int)ph.long_pairs.size() / 4;
  *desc_bytes_out = nb;
  // tight image with slack: a few reads land one column past the row end / on the row after the last (as in the reference)
  std::vector<uint8_t> tight((size_t)w * h + 64, 0);
  memcpy(tight.data(), img, (size_t)w * h);
  // S = (h+1) x (w+1) integral image, then one block per pixel: S(Y..Y+1, X..X+1) and the pixel I(Y-1, X+1) of the tight image
  std::vector<int32_t> S((size_t)(w + 1) * (h + 1), 0);
  for (int y = 0; y < h; ++y) { int s = 0; for (int x = 0; x < w; ++x) { s += img[(size_t)y * w + x]; S[(size_t)(y + 1) * (w + 1) + x + 1] = S[(size_t)y * (w + 1) + x + 1] + s; } }
  std::vector<Block4> blocks((size_t)w * h);
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      const int32_t* q = &S[(size_t)y * (w + 1) + x];
      blocks[(size_t)y * w + x] = encode_block(q[0], q[1], q[w + 1], q[w + 2], y ? tight[(size_t)(y - 1) * w + x + 1] : 0);
    }
  const BlockIntegral integ{blocks.data(), w};
  // describe_cull_kernel
  std::vector<KeyPoint> kept;
  std::vector<int> scales;
  for (int k = 0; k < n; ++k) {
    const KeyPoint kp = kps[k];
    int scale;
    if (scale_inv) {
      int lo = 0, hi = 63;
      while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (kp.size >= ph.scale_breaks[mid]) lo = mid; else hi = mid - 1; }
      scale = lo;
    } else scale = ph.basic_scale;
    const int border = (int)ph.size_list[scale];
    const float fb = (float)border, bx = (float)(w - border), by = (float)(h - border);
    if ((kp.x < fb) || (kp.x >= bx) || (kp.y < fb) || (kp.y >= by)) continue;
    kept.push_back(kp); scales.push_back(scale);
  }
  // describe_kernel
  std::vector<int> val(P);
  auto sample = [&](const KeyPoint& kp, int scale, int theta) {
    const float* pp = ph.points.data() + ((size_t)scale * 1024 + theta) * P * 3;
    for (int i = 0; i < P; ++i)
      val[i] = smoothed_intensity_t(tight.data(), w, integ, kp.x, kp.y, pp[3 * i], pp[3 * i + 1], pp[3 * i + 2],
                                    ph.sample_consts[((size_t)scale * P + i) * 2], ph.sample_consts[((size_t)scale * P + i) * 2 + 1]);
  };
  for (size_t k = 0; k < kept.size(); ++k) {
    KeyPoint& kp = kept[k];
    int theta = 0;
    if (rot_inv) {
      if (kp.angle == -1.0f) {
        sample(kp, scales[k], 0);
        int d0 = 0, d1 = 0;
        for (int p = 0; p < n_long; ++p) {
          const int* lp = &ph.long_pairs[4 * (size_t)p];
          const int delta = val[lp[0]] - val[lp[1]];
          d0 += delta * lp[2] / 1024;
          d1 += delta * lp[3] / 1024;
        }
        kp.angle = orientation_angle(d0, d1);
        theta = theta_from_estimated(kp.angle);
      } else theta = theta_from_given(kp.angle);
    }
    sample(kp, scales[k], theta);
    uint8_t* out = desc + k * nb;
    memset(out, 0, nb);
    for (int p = 0; p < n_short; ++p)
      if (val[ph.short_pairs[2 * (size_t)p]] > val[ph.short_pairs[2 * (size_t)p + 1]]) out[p >> 3] |= (uint8_t)(1u << (p & 7));
    kps[k] = kp;
  }
  return (int)kept.size();
}
