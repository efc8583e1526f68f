// Fused BRISK pyramid construction for a batch of frames (sm_100a).
//
// Replaces BriskScaleSpace::ConstructPyramid's chain of Halfsample8 /
// Twothirdsample8 calls (reference brisk/src/brisk-scale-space.cc:64-90,
// brisk/src/image-down-sampling.cc:142-392,550-787) by ONE kernel: each CTA
// pulls a 192x96 tile of layer 0 into shared memory with a TMA bulk tensor
// copy, derives the matching tiles of every other layer from it on chip
// (L1 = 2/3 L0, L2 = 1/2 L0, L3 = 1/2 L1, L4 = 1/2 L2, ...) and streams them
// out with 4-byte coalesced stores.  Layer 0 is read from HBM exactly once and
// nothing is re-read: traffic = |L0| + sum |Li| (+|L0| when layer 0 also has to
// be copied into the pyramid block).  The reference's per-column rounding
// regimes depend on the absolute column only, so tiles stay independent.
#include <cuda.h>
#include <cuda_runtime.h>

#include "brisk_math.cuh"
#include "kernels.h"

namespace briskb200 {

constexpr int kTileW = 192, kTileH = 96;  // layer-0 tile; 96 = 3 * 2^5 keeps 12 layers tile-local
constexpr int kPyrThreads = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// dst tile (dw x dh at dst layer origin (dx0, dy0)) = Halfsample8 of the src
// tile in shared memory.  src_w = full width of the source LAYER.
__device__ void half_tile(const uint8_t* __restrict__ s, int sw, int src_w, uint8_t* __restrict__ d, int dw, int dh, int dx0) {
  const int tid = threadIdx.x;
  if ((dw & 3) == 0) {
    const int groups = dw >> 2;
    const int hsize = src_w >> 4, body = (hsize >> 1) << 4, half_end = body + ((hsize & 1) ? 8 : 0);
    for (int i = tid; i < groups * dh; i += kPyrThreads) {
      const int oy = i / groups, g = i - oy * groups;
      const uint2 a = *reinterpret_cast<const uint2*>(s + (2 * oy) * sw + 8 * g);
      const uint2 b = *reinterpret_cast<const uint2*>(s + (2 * oy + 1) * sw + 8 * g);
      const int c = dx0 + 4 * g;  // absolute output column of the group's first pixel
      uint32_t out;
      if (c < half_end) {
        const uint32_t v0 = __vavgu4(a.x, b.x), v1 = __vavgu4(a.y, b.y);  // vertical pavgb
        const uint32_t e = __byte_perm(v0, v1, 0x6420), o = __byte_perm(v0, v1, 0x7531);
        out = (c < body) ? __vavgu4(e, o) : __vhaddu4(e, o);
      } else {
        out = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t lo = (k < 2) ? a.x : a.y, lo2 = (k < 2) ? b.x : b.y;
          const int sh = (k & 1) * 16;
          const int sum = ((lo >> sh) & 0xff) + ((lo >> (sh + 8)) & 0xff) + ((lo2 >> sh) & 0xff) + ((lo2 >> (sh + 8)) & 0xff);
          out |= (uint32_t)((sum + 2) >> 2) << (8 * k);
        }
      }
      *reinterpret_cast<uint32_t*>(d + oy * dw + 4 * g) = out;
    }
  } else {
    for (int i = tid; i < dw * dh; i += kPyrThreads) {
      const int oy = i / dw, ox = i - oy * dw;
      const uint8_t* a = s + (2 * oy) * sw + 2 * ox;
      d[oy * dw + ox] = (uint8_t)halfsample_px(a[0], a[1], a[sw], a[sw + 1], dx0 + ox, src_w);
    }
  }
}

// Horizontal step of Twothirdsample8's SSE body on four triples at once: w0..w2 hold twelve vertically
// combined bytes b0..b11; per triple (b0,b1,b2) the two outputs are avg(avg(b0,b1),b0) and avg(avg(b2,b1),b2)
// (pavgb).  Returns the eight output bytes in order.
__device__ __forceinline__ uint2 twothird_row4(uint32_t w0, uint32_t w1, uint32_t w2) {
  const uint32_t A = __byte_perm(__byte_perm(w0, w1, 0x0630), w2, 0x5210);  // b0 b3 b6 b9
  const uint32_t B = __byte_perm(__byte_perm(w0, w1, 0x0741), w2, 0x6210);  // b1 b4 b7 b10
  const uint32_t C = __byte_perm(__byte_perm(w0, w1, 0x0052), w2, 0x7410);  // b2 b5 b8 b11
  const uint32_t o0 = __vavgu4(__vavgu4(A, B), A), o1 = __vavgu4(__vavgu4(C, B), C);
  return make_uint2(__byte_perm(o0, o1, 0x5140), __byte_perm(o0, o1, 0x7362));
}

// dst tile = Twothirdsample8 of the 192x96 layer-0 tile; tx0 = absolute index
// of the tile's first source triple.  Four triples (12 source columns, 8 outputs per row) per step on packed
// bytes where all four lie in the reference's 15-column SSE blocks; the scalar tail columns go one 3x3
// block at a time.
__device__ void twothird_tile(const uint8_t* __restrict__ s, int src_w, uint8_t* __restrict__ d, int tx0) {
  constexpr int bw = kTileW / 3, bh = kTileH / 3, dw = 2 * bw, groups = bw / 4;
  const int sse_triples = 5 * (src_w / 15);
  for (int i = threadIdx.x; i < groups * bh; i += kPyrThreads) {
    const int by = i / groups, gx = i - by * groups;
    if (tx0 + 4 * gx + 3 < sse_triples) {
      const uint32_t* r0 = reinterpret_cast<const uint32_t*>(s + (3 * by) * kTileW + 12 * gx);
      const uint32_t* r1 = r0 + kTileW / 4;
      const uint32_t* r2 = r1 + kTileW / 4;
      uint32_t u[3], l[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const uint32_t a = r0[k], b = r1[k], c = r2[k];
        u[k] = __vavgu4(__vavgu4(a, b), a);  // upper output row: rows 0 and 1
        l[k] = __vavgu4(__vavgu4(c, b), c);  // lower output row: rows 2 and 1
      }
      *reinterpret_cast<uint2*>(d + (2 * by) * dw + 8 * gx) = twothird_row4(u[0], u[1], u[2]);
      *reinterpret_cast<uint2*>(d + (2 * by + 1) * dw + 8 * gx) = twothird_row4(l[0], l[1], l[2]);
    } else {
      for (int bx = 4 * gx; bx < 4 * gx + 4; ++bx) {
        int p[9], o[4];
#pragma unroll
        for (int k = 0; k < 9; ++k) p[k] = s[(3 * by + k / 3) * kTileW + 3 * bx + (k % 3)];
        twothird_block(p, tx0 + bx, src_w, o);
        *reinterpret_cast<uchar2*>(d + (2 * by) * dw + 2 * bx) = make_uchar2((uint8_t)o[0], (uint8_t)o[1]);
        *reinterpret_cast<uchar2*>(d + (2 * by + 1) * dw + 2 * bx) = make_uchar2((uint8_t)o[2], (uint8_t)o[3]);
      }
    }
  }
}

// Stream a finished tile from shared memory to its layer in HBM.
__device__ void flush_tile(const uint8_t* __restrict__ s, int tw, int th, uint8_t* __restrict__ layer, const LayerGeom& L,
                           int x0, int y0) {
  if ((tw & 15) == 0) {
    // 16-byte stores: tile origins and row pitches are multiples of 16, and writing the padding columns
    // [w, pitch) of a row is harmless
    const int groups = tw >> 4;
    for (int i = threadIdx.x; i < groups * th; i += kPyrThreads) {
      const int r = i / groups, g = i - r * groups;
      const int x = x0 + 16 * g, y = y0 + r;
      if (x < L.pitch && x < L.w && y < L.h) *reinterpret_cast<uint4*>(layer + (long long)y * L.pitch + x) = *reinterpret_cast<const uint4*>(s + r * tw + 16 * g);
    }
  } else if ((tw & 3) == 0) {
    const int groups = tw >> 2;
    for (int i = threadIdx.x; i < groups * th; i += kPyrThreads) {
      const int r = i / groups, g = i - r * groups;
      const int x = x0 + 4 * g, y = y0 + r;
      if (x < L.w && y < L.h) *reinterpret_cast<uint32_t*>(layer + (long long)y * L.pitch + x) = *reinterpret_cast<const uint32_t*>(s + r * tw + 4 * g);
    }
  } else {
    const int groups = tw >> 1;
    for (int i = threadIdx.x; i < groups * th; i += kPyrThreads) {
      const int r = i / groups, g = i - r * groups;
      const int x = x0 + 2 * g, y = y0 + r;
      if (x < L.w && y < L.h) *reinterpret_cast<uint16_t*>(layer + (long long)y * L.pitch + x) = *reinterpret_cast<const uint16_t*>(s + r * tw + 2 * g);
    }
  }
}

__global__ void __launch_bounds__(kPyrThreads)
pyramid_kernel(const __grid_constant__ CUtensorMap src_map, PyramidGeom g, uint8_t* __restrict__ pyr, int write_l0) {
  // tile buffers: [0] layer 0, then chain A (odd layers), chain B (even layers)
  __shared__ __align__(128) uint8_t s0[kTileW * kTileH];
  __shared__ __align__(16) uint8_t sa[2][128 * 64];  // ping-pong, chain A: 128x64, 64x32, ...
  __shared__ __align__(16) uint8_t sb[2][96 * 48];   // ping-pong, chain B: 96x48, 48x24, ...
  __shared__ __align__(8) uint64_t mbar;

  const int tx = blockIdx.x, ty = blockIdx.y, frame = blockIdx.z;
  uint8_t* fp = pyr + (long long)frame * g.frame_elems;

  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(kTileW * kTileH) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(s0)), "l"(&src_map), "r"(tx * kTileW), "r"(ty * kTileH), "r"(frame), "r"(smem_u32(&mbar))
        : "memory");
  }
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done) : "r"(smem_u32(&mbar)) : "memory");
    }
  }

  if (write_l0) {
    const LayerGeom& L = g.L[0];
    uint8_t* dst = fp + L.off;
    constexpr int groups = kTileW / 16;
    for (int i = threadIdx.x; i < groups * kTileH; i += kPyrThreads) {
      const int r = i / groups, c = i - r * groups;
      const int x = tx * kTileW + 16 * c, y = ty * kTileH + r;
      if (x < L.w && y < L.h) *reinterpret_cast<uint4*>(dst + (long long)y * L.pitch + x) = *reinterpret_cast<const uint4*>(s0 + r * kTileW + 16 * c);
    }
  }
  if (g.n_layers == 1) return;

  // level 1: L1 (2/3) and L2 (1/2) from the layer-0 tile
  twothird_tile(s0, g.w0, sa[0], tx * (kTileW / 3));
  if (g.n_layers > 2) half_tile(s0, kTileW, g.w0, sb[0], 96, 48, tx * 96);
  __syncthreads();
  flush_tile(sa[0], 128, 64, fp + g.L[1].off, g.L[1], tx * 128, ty * 64);
  if (g.n_layers > 2) flush_tile(sb[0], 96, 48, fp + g.L[2].off, g.L[2], tx * 96, ty * 48);

  // deeper levels: L(2k+1) = 1/2 L(2k-1), L(2k+2) = 1/2 L(2k)
  int aw = 128, ah = 64, bw = 96, bh = 48, cur = 0;
  for (int k = 1; 2 * k + 1 < g.n_layers; ++k) {
    const int la = 2 * k + 1, lb = 2 * k + 2;
    half_tile(sa[cur], aw, g.L[la - 2].w, sa[cur ^ 1], aw >> 1, ah >> 1, tx * (aw >> 1));
    if (lb < g.n_layers) half_tile(sb[cur], bw, g.L[lb - 2].w, sb[cur ^ 1], bw >> 1, bh >> 1, tx * (bw >> 1));
    __syncthreads();
    aw >>= 1; ah >>= 1; bw >>= 1; bh >>= 1; cur ^= 1;
    flush_tile(sa[cur], aw, ah, fp + g.L[la].off, g.L[la], tx * aw, ty * ah);
    if (lb < g.n_layers) flush_tile(sb[cur], bw, bh, fp + g.L[lb].off, g.L[lb], tx * bw, ty * bh);
  }
}

// ---------------------------------------------------------------------------
// 16-bit samplers: Halfsample16 / Twothirdsample16 (reference brisk/src/image-down-sampling.cc:56-139, 394-548).
// Stand-alone primitives (the reference's own 16-bit extractor path never gets this far: it integrates an empty image,
// SURVEY.md F10); HBM bound, one output pixel pair per thread, 32-bit coalesced stores.
// ---------------------------------------------------------------------------

// avg(avg(a, b), avg(sat(sat(c + 1) + 1), d)) with pavgw's rounding-up average (image-down-sampling.cc:117-122).
__device__ __forceinline__ uint32_t half16_px(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  c = min(min(c + 1u, 65535u) + 1u, 65535u);
  return (((a + b + 1u) >> 1) + ((c + d + 1u) >> 1) + 1u) >> 1;
}

__global__ void __launch_bounds__(256)
halfsample16_kernel(const uint16_t* __restrict__ src, int sw, int sh, long long spitch, uint16_t* __restrict__ dst, long long dpitch) {
  const int dw = sw / 2, dh = sh / 2;
  const int r = blockIdx.y;
  const int c = 2 * (blockIdx.x * blockDim.x + threadIdx.x);   // two output pixels per thread
  if (r >= dh || c >= dw) return;
  const uint16_t* s0 = src + (long long)(2 * r) * spitch + 2 * c;
  const uint16_t* s1 = s0 + spitch;
  uint16_t* d = dst + (long long)r * dpitch + c;
  d[0] = (uint16_t)half16_px(s0[0], s0[1], s1[0], s1[1]);
  if (c + 1 < dw) d[1] = (uint16_t)half16_px(s0[2], s0[3], s1[2], s1[3]);
}

// One thread per 3x3 source block -> 2x2 outputs: 4:2:2:1 weights, truncating division by 9, then the SIGNED
// saturation of packssdw (image-down-sampling.cc:503-523): sums above 32767 are stored as 32767.
__global__ void __launch_bounds__(256)
twothirdsample16_kernel(const uint16_t* __restrict__ src, int sw, int sh, long long spitch, uint16_t* __restrict__ dst, long long dpitch) {
  const int bw = sw / 3, bh = sh / 3;
  const int R = blockIdx.y, T = blockIdx.x * blockDim.x + threadIdx.x;
  if (R >= bh || T >= bw) return;
  uint32_t p[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) p[k] = src[(long long)(3 * R + k / 3) * spitch + 3 * T + k % 3];
  const uint32_t o0 = (4 * p[0] + 2 * p[1] + 2 * p[3] + p[4]) / 9, o1 = (4 * p[2] + 2 * p[1] + 2 * p[5] + p[4]) / 9;
  const uint32_t o2 = (4 * p[6] + 2 * p[7] + 2 * p[3] + p[4]) / 9, o3 = (4 * p[8] + 2 * p[7] + 2 * p[5] + p[4]) / 9;
  uint16_t* d = dst + (long long)(2 * R) * dpitch + 2 * T;
  d[0] = (uint16_t)min(o0, 32767u); d[1] = (uint16_t)min(o1, 32767u);
  d[dpitch] = (uint16_t)min(o2, 32767u); d[dpitch + 1] = (uint16_t)min(o3, 32767u);
}

cudaError_t launch_halfsample16(const uint16_t* src, int w, int h, long long spitch, uint16_t* dst, long long dpitch, cudaStream_t stream) {
  if (w / 2 <= 0 || h / 2 <= 0) return cudaSuccess;
  dim3 grid((w / 2 + 511) / 512, h / 2);
  halfsample16_kernel<<<grid, 256, 0, stream>>>(src, w, h, spitch, dst, dpitch);
  return cudaGetLastError();
}

cudaError_t launch_twothirdsample16(const uint16_t* src, int w, int h, long long spitch, uint16_t* dst, long long dpitch, cudaStream_t stream) {
  if (w / 3 <= 0 || h / 3 <= 0) return cudaSuccess;
  dim3 grid((w / 3 + 255) / 256, h / 3);
  twothirdsample16_kernel<<<grid, 256, 0, stream>>>(src, w, h, spitch, dst, dpitch);
  return cudaGetLastError();
}

cudaError_t launch_pyramid(const CUtensorMap& src_map, const PyramidGeom& g, uint8_t* pyr, int n_frames, int write_l0,
                           cudaStream_t stream) {
  dim3 grid((g.w0 + kTileW - 1) / kTileW, (g.h0 + kTileH - 1) / kTileH, n_frames);
  pyramid_kernel<<<grid, kPyrThreads, 0, stream>>>(src_map, g, pyr, write_l0);
  return cudaGetLastError();
}

}  // namespace briskb200
