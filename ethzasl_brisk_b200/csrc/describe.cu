// Integral image and BRISK descriptor extraction kernels (sm_100a).
//
// Replaces IntegralImage8 (reference brisk/include/brisk/internal/
// integral-image.h:56-161) and BriskDescriptorExtractor::doDescriptorComputation
// (brisk/src/brisk-descriptor-extractor.cc:612-778): border cull, orientation
// from the long pairs, rotated sampling, and bit packing of the short pairs.
// One warp per key point: lanes sample the 60/66 pattern points through the
// integral image, reduce the long-pair gradient with shuffles, and pack the
// comparison bits with __ballot_sync (bit p%32 of word p/32 is exactly lane
// order, reference :538-564).
#include <cuda_runtime.h>

#include <algorithm>

#include "describe_logic.cuh"
#include "kernels.h"

namespace briskb200 {

// --- integral image in one pass over the output ---
//
// S(y+1, x+1) = sum of I over rows <= y and columns <= x.  The frame is cut into vertical strips of 256
// columns.  Kernel 1 (one warp per row) writes, for every row and strip, the sum of the row's pixels LEFT of
// the strip.  Kernel 2 (one CTA per strip, one thread per column) walks down the rows, eight at a time,
// keeping the running column sums in registers; the eight rows of column sums go through shared memory to
// the eight warps, each of which turns one row into its prefix sums (eight consecutive columns per lane,
// then one warp scan) and adds what lies left of the strip.  The result is stored as one 16-byte BLOCK per pixel
// (describe_logic.cuh: the 2x2 neighbourhood S(Y..Y+1, X..X+1) plus the pixels I(Y, X) and I(Y-1, X+1), Y < h, X < w):
// the descriptor sampler then gets the twelve taps and the four corner pixels of a box from four 16-byte loads
// instead of sixteen scattered 4- and 1-byte loads (describe is bound by the L1 wavefront rate of exactly those gathers).  A thread
// builds its blocks from its own column, its left neighbour's (the sum left of the strip for the first thread)
// and the row before.  Traffic: the image twice (1 byte per pixel each) and 16 bytes per pixel out.
constexpr int kIntStrip = 256, kIntRows = 8;

__host__ __device__ inline int integral_strips(int w) { return (w + kIntStrip - 1) / kIntStrip; }
long long integral_aux_elems(int w, int h) { return (long long)h * integral_strips(w); }

__global__ void __launch_bounds__(256)
integral_left_sums_kernel(const uint8_t* __restrict__ imgs, long long frame_stride, int pitch, int w, int h, int32_t* __restrict__ aux) {
  const int frame = blockIdx.y, lane = threadIdx.x & 31;
  const int y = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (y >= h) return;
  const int ns = integral_strips(w);
  const uint8_t* row = imgs + (long long)frame * frame_stride + (long long)y * pitch;
  int32_t* out = aux + ((long long)frame * h + y) * ns;
  int left = 0;
  for (int s = 0; s < ns; ++s) {
    if (lane == 0) out[s] = left;
    // 256 bytes of the strip: two words per lane, bytes past the row end masked off (the pitch is a multiple of 16)
    int sum = 0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int x = s * kIntStrip + 4 * (lane + 32 * k);
      uint32_t v = 0;
      if (x < pitch) v = *reinterpret_cast<const uint32_t*>(row + x);
      if (x + 4 > w) v = x < w ? (v & (0xffffffffu >> (8 * (x + 4 - w)))) : 0u;
      sum += __vsadu4(v, 0u);
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    left += sum;
  }
}

__global__ void __launch_bounds__(kIntStrip)
integral_strip_kernel(const uint8_t* __restrict__ imgs, long long frame_stride, int pitch, int w, int h,
                      const int32_t* __restrict__ aux, int4* __restrict__ blocks) {
  __shared__ __align__(16) int s_t[2][kIntRows][kIntStrip];
  __shared__ int s_lf[2][kIntRows];   // left of the strip, rows <= y0 + r: S(y0 + r + 1, strip start)
  const int strip = blockIdx.x, frame = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ns = integral_strips(w);
  const int x = strip * kIntStrip + tid;
  const bool in = x < w;
  const uint8_t* col = imgs + (long long)frame * frame_stride + x;
  const int32_t* lft = aux + (long long)frame * h * ns + strip;
  int acc = 0, left_run = 0;
  int v[kIntRows], l[kIntRows];
  const uint8_t* cp = col;          // first row of the NEXT step's loads
  const int32_t* lp = lft;
  int4* op = blocks + (long long)frame * w * h + x;   // block(y0, x)
  int up_own = 0, up_left = 0;      // S(y0, x + 1) and S(y0, x): the row above the step (row 0 of S is zero)
  // I(Y - 1, x + 1) of the tightly packed image (the pixel right of the last column is the first of the next row), as
  // up_right[Y * pitch]; the neighbour thread loaded these bytes one step ago
  const uint8_t* up_right = imgs + (long long)frame * frame_stride + (x + 1 < w ? x + 1 - pitch : 0);
  const uint8_t* urp = up_right;   // up_right + y0 * pitch
#pragma unroll
  for (int r = 0; r < kIntRows; ++r) { v[r] = (in && r < h) ? cp[r * pitch] : 0; l[r] = r < h ? lp[r * ns] : 0; }
  for (int y0 = 0, it = 0; y0 < h; y0 += kIntRows, ++it) {
    int (*t)[kIntStrip] = s_t[it & 1];
    int left_mine = left_run;  // left of the strip, rows <= y0 + warp (the row this thread's warp will finish)
#pragma unroll
    for (int r = 0; r < kIntRows; ++r) {
      acc += v[r];             // column sum over rows <= y0 + r
      t[r][tid] = acc;
      left_run += l[r];
      if (tid == 0) s_lf[it & 1][r] = left_run;
      if (r <= warp) left_mine += l[r];
    }
    // next step's loads, in flight during the scans
    cp += kIntRows * pitch; lp += kIntRows * ns;
    if (y0 + 2 * kIntRows <= h) {
#pragma unroll
      for (int r = 0; r < kIntRows; ++r) { v[r] = in ? cp[r * pitch] : 0; l[r] = lp[r * ns]; }
    } else {
#pragma unroll
      for (int r = 0; r < kIntRows; ++r) {
        const bool row_in = y0 + kIntRows + r < h;
        v[r] = (in && row_in) ? cp[r * pitch] : 0;
        l[r] = row_in ? lp[r * ns] : 0;
      }
    }
    __syncthreads();
    int ur[kIntRows];
    if (y0 > 0 && y0 + kIntRows <= h) {
#pragma unroll
      for (int r = 0; r < kIntRows; ++r) ur[r] = in ? urp[r * pitch] : 0;
    } else {
#pragma unroll
      for (int r = 0; r < kIntRows; ++r) ur[r] = (in && y0 + r > 0 && y0 + r < h) ? up_right[(long long)(y0 + r) * pitch] : 0;
    }
    urp += kIntRows * pitch;
    {
      // warp `warp` finishes row y0 + warp: eight consecutive columns per lane, serial prefix, warp scan of the totals
      int4* p = reinterpret_cast<int4*>(&t[warp][8 * lane]);
      int4 a = p[0], b = p[1];
      a.y += a.x; a.z += a.y; a.w += a.z; b.x += a.w; b.y += b.x; b.z += b.y; b.w += b.z;
      int inc = b.w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += u; }
      const int base = left_mine + inc - b.w;
      a.x += base; a.y += base; a.z += base; a.w += base; b.x += base; b.y += base; b.z += base; b.w += base;
      p[0] = a; p[1] = b;
    }
    __syncthreads();
    // t[r][tid] = S(y0 + r + 1, x + 1); block(y0 + r, x) = {S(y0+r, x), S(y0+r, x+1), S(y0+r+1, x), S(y0+r+1, x+1)}
#pragma unroll
    for (int r = 0; r < kIntRows; ++r) {
      if (y0 + r >= h) break;
      const int own = t[r][tid];
      const int left = tid ? t[r][tid - 1] : s_lf[it & 1][r];
      if (in) {
        const Block4 e = encode_block(up_left, up_own, left, own, ur[r]);
        op[(long long)r * w] = make_int4(e.a, e.b, e.c, e.d);
      }
      up_own = own; up_left = left;
    }
    op += (long long)kIntRows * w;
  }
}

cudaError_t launch_integral(const uint8_t* imgs, long long frame_stride, int pitch, int w, int h, int n_frames,
                            int32_t* integral, cudaStream_t stream) {
  // the per-row "left of the strip" sums live behind the block images (integral_aux_elems per frame)
  int32_t* aux = integral + (long long)n_frames * w * h * 4;
  dim3 g1((h + 7) / 8, n_frames);
  integral_left_sums_kernel<<<g1, 256, 0, stream>>>(imgs, frame_stride, pitch, w, h, aux);
  dim3 g2(integral_strips(w), n_frames);
  integral_strip_kernel<<<g2, kIntStrip, 0, stream>>>(imgs, frame_stride, pitch, w, h, aux, reinterpret_cast<int4*>(integral));
  return cudaGetLastError();
}

// --- pitched planes -> tightly packed frames (row stride == width), one launch for a whole chunk ---

__global__ void __launch_bounds__(256)
copy_tight_kernel(const uint8_t* __restrict__ src, long long src_frame_stride, int src_pitch, int w, int h,
                  uint8_t* __restrict__ dst, long long dst_frame_stride) {
  const int frame = blockIdx.y;
  const uint8_t* s = src + (long long)frame * src_frame_stride;
  uint8_t* d = dst + (long long)frame * dst_frame_stride;
  const int words_per_row = (w + 3) / 4;   // source rows are 4-byte aligned (pitch is a multiple of 16)
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)words_per_row * h; i += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(i / words_per_row), x = 4 * (int)(i - (long long)y * words_per_row);
    const uint32_t v = *reinterpret_cast<const uint32_t*>(s + (long long)y * src_pitch + x);
    uint8_t* o = d + (long long)y * w + x;
#pragma unroll
    for (int b = 0; b < 4; ++b)
      if (x + b < w) o[b] = (uint8_t)(v >> (8 * b));
  }
}

cudaError_t launch_copy_tight(const uint8_t* src, long long src_frame_stride, int src_pitch, int w, int h, int n_frames,
                              uint8_t* dst, long long dst_frame_stride, cudaStream_t stream) {
  const long long words = (long long)((w + 3) / 4) * h;
  dim3 grid((unsigned)std::min<long long>((words + 255) / 256, 256), n_frames);
  copy_tight_kernel<<<grid, 256, 0, stream>>>(src, src_frame_stride, src_pitch, w, h, dst, dst_frame_stride);
  return cudaGetLastError();
}

// --- border cull: stable per-frame compaction of the key points that keep the
// whole pattern inside the image (reference :636-662) ---

__global__ void __launch_bounds__(256)
describe_cull_kernel(PatternDev pat, int w, int h, const KeyPoint* __restrict__ kps, int* __restrict__ counts, int kp_cap,
                     KeyPoint* __restrict__ kps_out, int* __restrict__ scales_out) {
  __shared__ int s_warp[8];
  __shared__ int s_base;
  const int frame = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // A list that did not fit into kp_cap slots is truncated, i.e. not the reference's key-point set: the frame keeps
  // its true count (> kp_cap), nothing of it is described and the call reports BRISK_ERR_CAPACITY.
  if (counts[frame] > kp_cap) return;
  const int n = counts[frame];
  const KeyPoint* src = kps + (long long)frame * kp_cap;
  KeyPoint* dst = kps_out + (long long)frame * kp_cap;
  int* sdst = scales_out + (long long)frame * kp_cap;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 256) {
    const int k = base + tid;
    bool keep = false;
    KeyPoint kp;
    int scale = 0;
    if (k < n) {
      kp = src[k];
      if (pat.scale_inv) {
        // scale index = #{s >= 1 : size >= break[s]}; the breaks are computed on
        // the host with the reference's own expression (:639-646), so the index
        // is exact by construction.  NaN / non-positive sizes map to 0.
        int lo = 0, hi = 63;
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if (kp.size >= pat.scale_breaks[mid]) lo = mid; else hi = mid - 1;
        }
        scale = lo;
      } else {
        scale = pat.basic_scale;
      }
      const int border = (int)pat.size_list[scale];
      const float fb = (float)border, bx = (float)(w - border), by = (float)(h - border);
      keep = !((kp.x < fb) || (kp.x >= bx) || (kp.y < fb) || (kp.y >= by));
    }
    const uint32_t m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    int before = s_base;
    for (int i = 0; i < warp; ++i) before += s_warp[i];
    const int pos = before + __popc(m & ((1u << lane) - 1));
    if (keep) { dst[pos] = kp; sdst[pos] = scale; }
    __syncthreads();
    if (tid == 0) { int t = 0; for (int i = 0; i < 8; ++i) t += s_warp[i]; s_base += t; }
    __syncthreads();
  }
  __syncthreads();
  if (tid == 0) counts[frame] = s_base;
}

// --- descriptor: one warp per key point ---

#ifndef BRISK_DESC_WARPS
#define BRISK_DESC_WARPS 8
#endif
constexpr int kDescWarps = BRISK_DESC_WARPS;
#ifndef BRISK_DESC_MIN_BLOCKS
#define BRISK_DESC_MIN_BLOCKS 5
#endif
constexpr int kMaxPoints = 96;

// All pattern points of one key point, spread over the warp.  The sampler is latency bound (16
// scattered loads per point), so every lane works on its first two points at once: the two
// independent gather chains overlap.
__device__ __forceinline__ void sample_pattern(const PatternDev& pat, const uint8_t* __restrict__ img, int pitch,
                                               const BlockIntegral integ, float kx, float ky,
                                               const float2* __restrict__ pp, int scale, int P, int lane, int* val) {
  const int4* cc = pat.sample_consts + scale * P;
  const int i0 = lane, i1 = lane + 32;
  if (i1 < P) {
    const int4 c0 = cc[i0], c1 = cc[i1];
    const float2 p0 = pp[i0], p1 = pp[i1];
    const int v0 = smoothed_intensity_t(img, pitch, integ, kx, ky, p0.x, p0.y, __int_as_float(c0.z), c0.x, c0.y);
    const int v1 = smoothed_intensity_t(img, pitch, integ, kx, ky, p1.x, p1.y, __int_as_float(c1.z), c1.x, c1.y);
    val[i0] = v0; val[i1] = v1;
  } else if (i0 < P) {
    const int4 c0 = cc[i0];
    const float2 p0 = pp[i0];
    val[i0] = smoothed_intensity_t(img, pitch, integ, kx, ky, p0.x, p0.y, __int_as_float(c0.z), c0.x, c0.y);
  }
  for (int i = lane + 64; i < P; i += 32) {
    const int4 sc = cc[i];
    const float2 pt = pp[i];
    val[i] = smoothed_intensity_t(img, pitch, integ, kx, ky, pt.x, pt.y, __int_as_float(sc.z), sc.x, sc.y);
  }
}

__global__ void __launch_bounds__(kDescWarps * 32, BRISK_DESC_MIN_BLOCKS)
describe_kernel(PatternDev pat, const uint8_t* __restrict__ imgs, long long frame_stride, int pitch, int w, int h,
                const int32_t* __restrict__ integral, const KeyPoint* __restrict__ kps_in, const int* __restrict__ scales,
                const int* __restrict__ counts, int kp_cap, KeyPoint* __restrict__ kps_out, uint8_t* __restrict__ desc) {
  __shared__ int s_val[kDescWarps][kMaxPoints];
  const int frame = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k = blockIdx.x * kDescWarps + warp;
  if (k >= counts[frame] || counts[frame] > kp_cap) return;
  const long long slot = (long long)frame * kp_cap + k;
  const KeyPoint kp = kps_in[slot];
  const int scale = scales[slot];
  const uint8_t* img = imgs + (long long)frame * frame_stride;
  const BlockIntegral integ{reinterpret_cast<const Block4*>(integral) + (long long)frame * w * h, w};
  const int P = pat.n_points;
  int* val = s_val[warp];

  int theta = 0;
  float angle = kp.angle;
  if (pat.rot_inv) {
    if (kp.angle == -1.0f) {
      // un-rotated samples, long-pair gradient (:697-739)
      const float2* pp = pat.points + ((long long)scale * 1024) * P;
      sample_pattern(pat, img, pitch, integ, kp.x, kp.y, pp, scale, P, lane, val);
      __syncwarp();
      int d0 = 0, d1 = 0;
      for (int p = lane; p < pat.n_long; p += 32) {
        const int2 lp = pat.long_pairs[p];
        const int delta = val[lp.x & 0xffff] - val[lp.x >> 16];
        d0 += delta * (int)(short)(lp.y & 0xffff) / 1024;
        d1 += delta * (lp.y >> 16) / 1024;
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) { d0 += __shfl_xor_sync(0xffffffffu, d0, o); d1 += __shfl_xor_sync(0xffffffffu, d1, o); }
      angle = orientation_angle(d0, d1);
      theta = theta_from_estimated(angle);
      __syncwarp();
    } else {
      theta = theta_from_given(kp.angle);
    }
  }
  // samples in the rotated pattern (:755-772)
  const float2* pp = pat.points + ((long long)scale * 1024 + theta) * P;
  sample_pattern(pat, img, pitch, integ, kp.x, kp.y, pp, scale, P, lane, val);
  __syncwarp();
  // short-pair comparisons -> bits (:538-564); rows are zero-padded to desc_bytes
  uint32_t* out = reinterpret_cast<uint32_t*>(desc + slot * pat.desc_bytes);
  const int words = pat.desc_bytes >> 2;
  for (int wd = 0; wd < words; ++wd) {
    const int p = wd * 32 + lane;
    bool bit = false;
    if (p < pat.n_short) {
      const ushort2 sp = *reinterpret_cast<const ushort2*>(pat.short_pairs + 2 * p);
      bit = val[sp.x] > val[sp.y];
    }
    const uint32_t word = __ballot_sync(0xffffffffu, bit);
    if (lane == 0) out[wd] = word;
  }
  if (lane == 0) {
    KeyPoint o = kp;
    o.angle = angle;
    kps_out[slot] = o;
  }
}

cudaError_t launch_describe(const PatternDev& pat, const uint8_t* imgs, long long frame_stride, int pitch, int w, int h,
                            int n_frames, const int32_t* integral, KeyPoint* kps, int* counts, int kp_cap,
                            KeyPoint* kps_scratch, int* scale_scratch, uint8_t* desc, cudaStream_t stream) {
  describe_cull_kernel<<<n_frames, 256, 0, stream>>>(pat, w, h, kps, counts, kp_cap, kps_scratch, scale_scratch);
  dim3 grid((kp_cap + kDescWarps - 1) / kDescWarps, n_frames);
  describe_kernel<<<grid, kDescWarps * 32, 0, stream>>>(pat, imgs, frame_stride, pitch, w, h, integral, kps_scratch,
                                                        scale_scratch, counts, kp_cap, kps, desc);
  return cudaGetLastError();
}

}  // namespace briskb200
