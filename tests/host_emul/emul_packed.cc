// TEST HARNESS: the product's packed-lane device code (ethzasl_brisk_b200/csrc/fast_packed.cuh: two pixels per register on
// 16-bit lanes) compiled for the host, with the handful of SIMD intrinsics it uses restated in plain C++ -- so that the
// packed FAST score rows, the packed 9-of-16 test of the detector's second phase and its compass pre-test can be checked
// against the scalar logic (brisk_math.cuh) without a GPU.  Not linked by the product.
#include <stdint.h>
#include <string.h>
#include <algorithm>

// --- CUDA intrinsics on two unsigned 16-bit lanes / bytes, as the PTX ISA defines them ---
static inline uint32_t lanes(uint32_t a, uint32_t b, uint32_t c, bool want_max) {
  uint32_t r = 0;
  for (int s = 0; s < 32; s += 16) {
    const uint32_t x = (a >> s) & 0xffffu, y = (b >> s) & 0xffffu, z = (c >> s) & 0xffffu;
    const uint32_t m = want_max ? std::max(x, std::max(y, z)) : std::min(x, std::min(y, z));
    r |= m << s;
  }
  return r;
}
static inline uint32_t __vimax3_u16x2(uint32_t a, uint32_t b, uint32_t c) { return lanes(a, b, c, true); }
static inline uint32_t __vimin3_u16x2(uint32_t a, uint32_t b, uint32_t c) { return lanes(a, b, c, false); }
static inline uint32_t __vmaxu2(uint32_t a, uint32_t b) { return lanes(a, b, b, true); }
static inline uint32_t __vminu2(uint32_t a, uint32_t b) { return lanes(a, b, b, false); }
static inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t sel) {
  const uint64_t ab = ((uint64_t)b << 32) | a;
  uint32_t r = 0;
  for (int i = 0; i < 4; ++i) r |= (uint32_t)((ab >> (8 * ((sel >> (4 * i)) & 7))) & 0xffu) << (8 * i);   // (selectors 0..7 only)
  return r;
}
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t shift) {
  const uint64_t v = ((uint64_t)hi << 32) | lo;
  return (uint32_t)(v >> (shift & 31));
}
using std::max;
using std::min;

#define __CUDACC__ 1          // take the device-only section of fast_packed.cuh ...
#define __device__            // ... as plain host functions
#define __forceinline__ inline
#define __host__
#define __restrict__
#undef BRISK_HD
#include "../../ethzasl_brisk_b200/csrc/brisk_common.cuh"
#undef BRISK_HD
#define BRISK_HD inline
#include "../../ethzasl_brisk_b200/csrc/fast_packed.cuh"

using namespace briskb200;

// Dense FAST 9-16 scores (clipped at 0) of the interior of an image through fast916_row<NP>, NP = 1..3.
extern "C" int emul_packed_scores(const uint8_t* img, int w, int h, int pitch, int np, uint8_t* out) {
  for (int y = 3; y < h - 3; ++y)
    for (int x = 3; x < w - 3; x += 2 * np) {
      uint32_t f[3] = {0, 0, 0};
      if (np == 1) fast916_row<1>(img, pitch, h, x, y, f);
      else if (np == 2) fast916_row<2>(img, pitch, h, x, y, f);
      else fast916_row<3>(img, pitch, h, x, y, f);
      for (int j = 0; j < 2 * np && x + j < w - 3; ++j) out[(size_t)y * w + x + j] = (uint8_t)((f[j >> 1] >> ((j & 1) * 16)) & 0xffu);
    }
  return 0;
}

// The detector's phase 2 on pairs (detect.cu segment_test_pair): corner verdicts of two pixels at contrast b0 / b1 from the
// packed arc extrema.  ring0 / ring1: the 16 ring pixels, c0 / c1 the centres.  Returns bit 0 / bit 1.
extern "C" int emul_packed_segment_pair(const uint8_t* ring0, int c0, int b0, const uint8_t* ring1, int c1, int b1) {
  uint32_t p[16];
  for (int i = 0; i < 16; ++i) p[i] = (uint32_t)ring0[i] | ((uint32_t)ring1[i] << 16);
  uint32_t bright, dark;
  arc_extrema16x2(p, &bright, &dark);
  constexpr uint32_t kHi = 0x80008000u;
  const uint32_t c = (uint32_t)c0 | ((uint32_t)c1 << 16), b = (uint32_t)b0 | ((uint32_t)b1 << 16);
  const uint32_t tb = ((c + b) | kHi) - bright, td = ((dark + b) | kHi) - c;
  const uint32_t hit = ~(tb & td) & kHi;
  return (int)((hit >> 15 & 1u) | (hit >> 30 & 2u));
}

// The detector's compass pre-test on pairs (detect.cu, row walk): candidate flags of two pixels from their four compass
// pixels (left, up, right, down), centre and contrast.
extern "C" int emul_packed_compass_pair(const uint8_t* compass0, int c0, int b0, const uint8_t* compass1, int c1, int b1) {
  auto pk = [&](int i) { return (uint32_t)compass0[i] | ((uint32_t)compass1[i] << 16); };
  const uint32_t p0 = pk(0), p4 = pk(1), p8 = pk(2), p12 = pk(3);
  const uint32_t c = (uint32_t)c0 | ((uint32_t)c1 << 16), b = (uint32_t)b0 | ((uint32_t)b1 << 16);
  constexpr uint32_t kHi = 0x80008000u;
  const uint32_t S2 = __vminu2(__vmaxu2(p0, p8), __vmaxu2(p4, p12)), s2 = __vmaxu2(__vminu2(p0, p8), __vminu2(p4, p12));
  const uint32_t bright = ((c + b) | kHi) - S2, dark = ((s2 + b) | kHi) - c;
  const uint32_t cand = ~(bright & dark) & kHi;
  return (int)((cand >> 15 & 1u) | (cand >> 30 & 2u));
}

// Descriptor byte -> eight E2M1 nibbles (brisk_math.cuh, the operand expansion of the FP4 matcher).
extern "C" uint32_t emul_e2m1_expand_byte(uint32_t byte) { return e2m1_expand_byte(byte); }
