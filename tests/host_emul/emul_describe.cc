// TEST HARNESS: the product's descriptor sampling logic (describe_logic.cuh)
// compiled for the host, run serially over key points.  Pattern tables come
// from the product's host-side builder (pattern.cc).  Not linked by the product.
#include <stdint.h>
#include <string.h>
#include <vector>

#include "../../ethzasl_brisk_b200/csrc/describe_logic.cuh"
#include "../../ethzasl_brisk_b200/csrc/pattern.h"

using namespace briskb200;

// samples: [n][P] values of the un-rotated pattern; kps must already be culled (all inside).
extern "C" int emul_samples(const uint8_t* img, int w, int h, const KeyPoint* kps, const int* scales, int n, int version,
                            float pattern_scale, int theta, int* samples) {
  static PatternHost ph;
  static int have_v = -1; static float have_ps = -1;
  if (have_v != version || have_ps != pattern_scale) { if (!build_pattern(version, pattern_scale, nullptr, &ph).empty()) return -1; have_v = version; have_ps = pattern_scale; }
  std::vector<int32_t> integ((size_t)(w + 1) * (h + 1), 0);
  for (int y = 0; y < h; ++y) { int s = 0; for (int x = 0; x < w; ++x) { s += img[(size_t)y * w + x]; integ[(size_t)(y + 1) * (w + 1) + x + 1] = integ[(size_t)y * (w + 1) + x + 1] + s; } }
  const int P = ph.n_points;
  for (int k = 0; k < n; ++k) {
    const float* pp = ph.points.data() + ((size_t)scales[k] * 1024 + theta) * P * 3;
    for (int i = 0; i < P; ++i)
      samples[(size_t)k * P + i] = smoothed_intensity(img, w, integ.data(), w + 1, kps[k].x, kps[k].y, pp[3 * i], pp[3 * i + 1], pp[3 * i + 2],
                                                      ph.sample_consts[((size_t)scales[k] * P + i) * 2], ph.sample_consts[((size_t)scales[k] * P + i) * 2 + 1]);
  }
  return P;
}
