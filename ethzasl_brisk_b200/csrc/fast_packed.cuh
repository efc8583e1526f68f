// FAST 9-16 scores of a run of horizontally adjacent pixels on packed pairs (device only, sm_100a).
//
// The per-corner NMS kernels look scores up in small dense blocks (the 5x5 window around a corner, the 4x4 scan
// tiles of the neighbouring layers).  Evaluating those pixel by pixel costs 16 scattered byte loads and ~70
// three-input min / max each, although the rings of neighbouring pixels overlap almost completely.  Here one call
// scores 2 * NP adjacent pixels of a row: seven row segments of 16 bytes are loaded as words, and every arithmetic
// instruction works on a PAIR of pixels (two 16-bit lanes, VIMNMX3.U16x2) -- per score about 7 loads and 50
// instructions instead of 16 loads and 150.  Same closed form as fast916 / arc_contrast16 (brisk_math.cuh):
// F = max over the 9-arcs of max(min(p) - c, c - max(p)), minus 1, clipped at 0.
#pragma once
#include "brisk_math.cuh"

namespace briskb200 {

#ifdef __CUDACC__

// (b_k, b_k+1) of the 12-byte row segment (words w0 w1 w2; h0 = bytes 2..5, h1 = bytes 6..9) as two 16-bit lanes.
// k is a compile-time constant after unrolling.
__device__ __forceinline__ uint32_t packed_pair_at(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t h0, uint32_t h1, int k) {
  switch (k) {
    case 0: return __byte_perm(w0, 0, 0x4140);
    case 1: return __byte_perm(w0, 0, 0x4241);
    case 2: return __byte_perm(w0, 0, 0x4342);
    case 3: return __byte_perm(h0, 0, 0x4241);
    case 4: return __byte_perm(w1, 0, 0x4140);
    case 5: return __byte_perm(w1, 0, 0x4241);
    case 6: return __byte_perm(w1, 0, 0x4342);
    case 7: return __byte_perm(h1, 0, 0x4241);
    case 8: return __byte_perm(w2, 0, 0x4140);
    case 9: return __byte_perm(w2, 0, 0x4241);
    default: return __byte_perm(w2, 0, 0x4342);
  }
}

// The two extrema that decide a 9-of-16 segment test, on PAIRS of pixels (p[i] = ring pixel i of two pixels in the two
// 16-bit lanes): *bright = max over the sixteen 9-arcs of the arc's minimum, *dark = min over the arcs of the arc's maximum.
// A pixel is a corner at contrast b  <=>  bright > c + b or dark < c - b; its score is max(bright - c, c - dark) - 1.
// Sliding min / max over 9 as windows of 3, then three of those 3 apart.
__device__ __forceinline__ void arc_extrema16x2(const uint32_t (&p)[16], uint32_t* bright_out, uint32_t* dark_out) {
  uint32_t lo3[16], hi3[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    lo3[i] = __vimin3_u16x2(p[i], p[(i + 1) & 15], p[(i + 2) & 15]);
    hi3[i] = __vimax3_u16x2(p[i], p[(i + 1) & 15], p[(i + 2) & 15]);
  }
  uint32_t lo9[16], hi9[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    lo9[i] = __vimin3_u16x2(lo3[i], lo3[(i + 3) & 15], lo3[(i + 6) & 15]);
    hi9[i] = __vimax3_u16x2(hi3[i], hi3[(i + 3) & 15], hi3[(i + 6) & 15]);
  }
  uint32_t bright = lo9[15], dark = hi9[15];
#pragma unroll
  for (int i = 0; i < 15; i += 3) {
    bright = __vmaxu2(bright, __vimax3_u16x2(lo9[i], lo9[i + 1], lo9[i + 2]));
    dark = __vminu2(dark, __vimin3_u16x2(hi9[i], hi9[i + 1], hi9[i + 2]));
  }
  *bright_out = bright; *dark_out = dark;
}

// F (clipped at 0) of the pixels (x + 2q, y) / (x + 2q + 1, y) in the low / high half of out[q], q < NP <= 3.
// Reads rows y-3..y+3 and columns x-3..x+2NP+2; rows and words outside the plane are clamped, which only changes
// the scores of pixels within 3 pixels of the image border -- the callers never use those (in_border).
template <int NP>
__device__ __forceinline__ void fast916_row(const uint8_t* __restrict__ img, int pitch, int h, int x, int y, uint32_t out[NP]) {
  const int cb = x - 3;               // column of byte 0 of the segment
  const int ab = cb & ~3;             // aligned word holding it
  const unsigned sh = (unsigned)(cb - ab) * 8u;
  int wx[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) wx[j] = min(max(ab + 4 * j, 0), pitch - 4);
  uint32_t p[NP][16], c[NP];
#pragma unroll
  for (int dy = -3; dy <= 3; ++dy) {
    const uint8_t* row = img + (long long)min(max(y + dy, 0), h - 1) * pitch;
    const uint32_t r0 = *reinterpret_cast<const uint32_t*>(row + wx[0]), r1 = *reinterpret_cast<const uint32_t*>(row + wx[1]);
    const uint32_t r2 = *reinterpret_cast<const uint32_t*>(row + wx[2]), r3 = *reinterpret_cast<const uint32_t*>(row + wx[3]);
    const uint32_t w0 = __funnelshift_r(r0, r1, sh), w1 = __funnelshift_r(r1, r2, sh), w2 = __funnelshift_r(r2, r3, sh);
    const uint32_t h0 = __funnelshift_r(w0, w1, 16), h1 = __funnelshift_r(w1, w2, 16);
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      // ring of agast/include/agast/oast9-16.h:99-116 (index -> dx, dy as in fast916); byte of dx for pair q: dx + 3 + 2q
#define BRISK_RP(i, dx) p[q][i] = packed_pair_at(w0, w1, w2, h0, h1, (dx) + 3 + 2 * q)
      if (dy == -3) { BRISK_RP(3, -1); BRISK_RP(4, 0); BRISK_RP(5, 1); }
      if (dy == -2) { BRISK_RP(2, -2); BRISK_RP(6, 2); }
      if (dy == -1) { BRISK_RP(1, -3); BRISK_RP(7, 3); }
      if (dy == 0) { BRISK_RP(0, -3); BRISK_RP(8, 3); c[q] = packed_pair_at(w0, w1, w2, h0, h1, 3 + 2 * q); }
      if (dy == 1) { BRISK_RP(15, -3); BRISK_RP(9, 3); }
      if (dy == 2) { BRISK_RP(14, -2); BRISK_RP(10, 2); }
      if (dy == 3) { BRISK_RP(13, -1); BRISK_RP(12, 0); BRISK_RP(11, 1); }
#undef BRISK_RP
    }
  }
#pragma unroll
  for (int q = 0; q < NP; ++q) {
    uint32_t bright, dark;
    arc_extrema16x2(p[q], &bright, &dark);
    // per lane max(bright - c, c - dark) - 1, clipped at 0; the lanes are biased by 0x8000 so that no borrow crosses them
    constexpr uint32_t kBias = 0x80008000u;
    const uint32_t d1 = (bright | kBias) - c[q], d2 = (c[q] | kBias) - dark;
    const uint32_t m = __vmaxu2(d1, d2);
    out[q] = __vmaxu2(m, kBias + 0x00010001u) - (kBias + 0x00010001u);
  }
}

#endif  // __CUDACC__

}  // namespace briskb200
