// Scale-space NMS and refinement kernels of the AGAST path (sm_100a); the
// per-corner logic lives in nms_logic.cuh (shared with the host-side tests).
//
// Replaces BriskScaleSpace::GetKeypoints and friends (reference
// brisk/src/brisk-scale-space.cc:92-1364).  Launch sequence per batch:
//   nms_prefix_kernel  one thread per raw corner   IsMax2D's 8 comparisons
//   nms_checks_kernel  one thread per raw corner   above / below scale checks (pure)
//   nms_chain_kernel   one CTA per frame           layer by layer: tying corners in
//                                                  raster order, then the footprint
//                                                  the accepted corners leave on the
//                                                  layer above
//   refine_kernel      one thread per raw corner   sub-pixel / scale refinement
//   compact_kernel     one CTA per frame           ordered compaction (+ mask filter)
#include <cuda_runtime.h>

#include "kernels.h"
#include "nms_logic.cuh"

namespace briskb200 {

struct FrameViews {
  LayerView v[kMaxLayers];
};

__device__ __forceinline__ void make_views(const PyramidGeom& g, const DetectWorkspace& ws, int frame, FrameViews* fv) {
  const long long fo = (long long)frame * g.frame_elems;
#pragma unroll 1
  for (int l = 0; l < g.n_layers; ++l) {
    const LayerGeom& L = g.L[l];
    fv->v[l] = LayerView{ws.pyr + fo + L.off, ws.cm + fo + L.off, ws.bm + fo + L.off, L.w, L.h, L.pitch, L.scale, L.offset};
  }
}

__device__ __forceinline__ void unpack_corner(uint32_t c, int* x, int* y, int* layer) {
  *x = c & 0x1fff; *y = (c >> 13) & 0x1fff; *layer = c >> 26;
}

__global__ void __launch_bounds__(128)
nms_prefix_kernel(PyramidGeom g, DetectWorkspace ws) {
  const int frame = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = min(ws.layer_start[(long long)frame * (kMaxLayers + 1) + g.n_layers], ws.corner_cap);
  if (k >= n) return;
  int x, y, layer;
  unpack_corner(ws.corners[(long long)frame * ws.corner_cap + k], &x, &y, &layer);
  const long long fo = (long long)frame * g.frame_elems;
  const LayerGeom& L = g.L[layer];
  const LayerView v{ws.pyr + fo + L.off, ws.cm + fo + L.off, ws.bm + fo + L.off, L.w, L.h, L.pitch, L.scale, L.offset};
  uint8_t fwin[25];
  nms_prefix(v, x, y, fwin);
  if (v.cm[(long long)y * L.pitch + x] & kCmTie) {
    uint8_t* dst = ws.fwin + ((long long)frame * ws.corner_cap + k) * 32;
#pragma unroll
    for (int i = 0; i < 25; ++i) dst[i] = fwin[i];
  }
}

__global__ void __launch_bounds__(128)
nms_checks_kernel(PyramidGeom g, DetectWorkspace ws) {
  const int frame = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = min(ws.layer_start[(long long)frame * (kMaxLayers + 1) + g.n_layers], ws.corner_cap);
  if (k >= n) return;
  int x, y, layer;
  unpack_corner(ws.corners[(long long)frame * ws.corner_cap + k], &x, &y, &layer);
  FrameViews fv;
  make_views(g, ws, frame, &fv);
  uint16_t* e = fv.v[layer].cm + (long long)y * fv.v[layer].pitch + x;
  const uint16_t ev = *e;
  if ((ev & kCmDecided) && !(ev & kCmAccept)) return;
  CheckResult r;
  const bool ok = nms_checks(fv.v, g.n_layers, layer, x, y, &r);
  if (ok) *e = ev | kCmChecks;
  // kept even when the checks fail: the footprint of the scan of the layer above is needed by the chain kernel
  *reinterpret_cast<CheckResult*>(ws.checks + ((long long)frame * ws.corner_cap + k) * 8) = r;
}

// Warp-cooperative IsMax2D tie path for one tying corner: the 64 corner-map entries of the 8x8
// window are staged by the lanes (two each), lanes 0..24 each reconstruct the value of one pixel of
// the 5x5 neighbourhood, and the 3x3 binomial sums around the centre and around every tying
// neighbour are formed with shuffles.  Returns 1 accept, 0 reject, -1 blocked (an earlier tying
// corner in the window is undecided); uniform across the warp.
__device__ __forceinline__ int warp_tie_decide(const LayerView& L, int mode, int x, int y, const uint8_t* __restrict__ fwin,
                                               uint16_t* s_win /* 64 entries, this warp's */) {
  const int lane = threadIdx.x & 31;
  bool blocked = false;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int i = lane + 32 * h;
    const int px = x - 4 + (i & 7), py = y - 4 + (i >> 3);
    uint16_t e = 0;
    if (py >= 3 && px >= 3 && px < L.w - 3 && py < L.h - 3) e = L.cm[(long long)py * L.pitch + px];
    s_win[i] = e;
    if ((e & kCmT) && !(e & kCmDecided) && (py < y || (py == y && px < x))) blocked = true;
  }
  if (__any_sync(0xffffffffu, blocked)) return -1;
  __syncwarp();
  const TieWindow W{s_win, 1, x - 4, y - 4};
  const int center = W.at(x, y) & kCmT;
  const int ox = lane % 5 - 2, oy = lane / 5 - 2;  // lanes 0..24 <-> 5x5 offsets, row-major
  int v = 0;
  if (lane < 25) v = tie_pixel_value(L, W, mode, x, y, ox, oy, fwin[lane], center);
  // binomial 3x3 sum centred on every lane's own pixel (meaningful for the inner 3x3 lanes)
  int sum = 0;
#pragma unroll
  for (int wy = -1; wy <= 1; ++wy)
#pragma unroll
    for (int wx = -1; wx <= 1; ++wx) {
      const int src = lane + wy * 5 + wx;
      const int t = __shfl_sync(0xffffffffu, v, src & 31);
      sum += ((wx == 0 ? 2 : 1) * (wy == 0 ? 2 : 1)) * t;
    }
  const int smoothed = __shfl_sync(0xffffffffu, sum, 12);
  const bool inner = lane < 25 && ox >= -1 && ox <= 1 && oy >= -1 && oy <= 1 && lane != 12;
  const bool beaten = inner && v == center && sum > smoothed;
  const int verdict = __any_sync(0xffffffffu, beaten) ? 0 : 1;
  __syncwarp();
  return verdict;
}

// One CTA per frame, layers in order (layer i+1 needs the touch marks that layer i's accepted
// corners leave on it).  Within a layer the tying corners are resolved in parallel rounds, one warp
// per corner: a corner whose raster-earlier tying neighbours are all decided is decidable, whatever
// the order.
constexpr int kChainThreads = 1024;
__global__ void __launch_bounds__(kChainThreads)
nms_chain_kernel(PyramidGeom g, DetectWorkspace ws, int* __restrict__ error_flag) {
  __shared__ FrameViews fv;
  __shared__ int s_left;
  __shared__ uint16_t s_win[kChainThreads / 32][64];
  const int frame = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) make_views(g, ws, frame, &fv);
  __syncthreads();
  const int* ls = ws.layer_start + (long long)frame * (kMaxLayers + 1);
  const uint32_t* corners = ws.corners + (long long)frame * ws.corner_cap;
  for (int layer = 0; layer < g.n_layers; ++layer) {
    const int mode = g.n_layers == 1 ? kModeSingle : (layer == g.n_layers - 1 ? kModeLast : kModeMid);
    const int begin = min(ls[layer], ws.corner_cap), end = min(ls[layer + 1], ws.corner_cap);
    const LayerView& L = fv.v[layer];
    // collect this layer's tying (undecided) corners; the key-point scratch of the frame is free until
    // refine_kernel runs and serves as the list
    int* tie_list = reinterpret_cast<int*>(ws.kp_tmp + (long long)frame * ws.corner_cap);
    if (tid == 0) s_left = 0;
    __syncthreads();
    for (int k = begin + tid; k < end; k += blockDim.x) {
      int x, y, l2;
      unpack_corner(corners[k], &x, &y, &l2);
      if (!(L.cm[(long long)y * L.pitch + x] & kCmDecided)) tie_list[atomicAdd(&s_left, 1)] = k;
    }
    __syncthreads();
    const int n_ties = s_left;
    __syncthreads();
    for (int round = 0; n_ties > 0; ++round) {
      if (tid == 0) s_left = 0;
      __syncthreads();
      int left = 0;
      for (int i = warp; i < n_ties; i += kChainThreads / 32) {
        const int k = tie_list[i];
        if (k < 0) continue;  // uniform across the warp
        int x, y, l2;
        unpack_corner(corners[k], &x, &y, &l2);
        const int verdict = warp_tie_decide(L, mode, x, y, ws.fwin + ((long long)frame * ws.corner_cap + k) * 32, s_win[warp]);
        if (verdict < 0) { ++left; continue; }
        if (lane == 0) {
          uint16_t* e = L.cm + (long long)y * L.pitch + x;
          *e = *e | (uint16_t)(kCmDecided | (verdict ? kCmAccept : 0));
          tie_list[i] = -1;
        }
        __syncwarp();
      }
      if (lane == 0 && left) atomicAdd(&s_left, left);
      __syncthreads();
      const int remaining = s_left;
      __syncthreads();
      if (remaining == 0) { if (tid == 0) ws.rounds[frame * kMaxLayers + layer] = round + 1; break; }
      if (round > (1 << 16)) { if (tid == 0) atomicExch(error_flag, 2); break; }  // cannot happen: dependencies are acyclic
    }
    // footprint of the accepted corners on the layer above
    if (mode == kModeMid) {
      for (int k = begin + tid; k < end; k += blockDim.x) {
        int x, y, l2;
        unpack_corner(corners[k], &x, &y, &l2);
        if (L.cm[(long long)y * L.pitch + x] & kCmAccept)
          mark_above(fv.v, layer, x, y, *reinterpret_cast<const CheckResult*>(ws.checks + ((long long)frame * ws.corner_cap + k) * 8));
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(128)
refine_kernel(PyramidGeom g, DetectWorkspace ws) {
  const int frame = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = min(ws.layer_start[(long long)frame * (kMaxLayers + 1) + g.n_layers], ws.corner_cap);
  if (k >= n) return;
  const long long slot = (long long)frame * ws.corner_cap + k;
  int x, y, layer;
  unpack_corner(ws.corners[slot], &x, &y, &layer);
  FrameViews fv;
  make_views(g, ws, frame, &fv);
  const uint16_t e = fv.v[layer].cm[(long long)y * fv.v[layer].pitch + x];
  bool valid = false;
  if ((e & kCmAccept) && (e & kCmChecks)) {
    const CheckResult r = *reinterpret_cast<const CheckResult*>(ws.checks + slot * 8);
    KeyPoint kp;
    valid = refine_emit(fv.v, g.n_layers, layer, x, y, r, &kp);
    if (valid) ws.kp_tmp[slot] = kp;
  }
  ws.kp_valid[slot] = valid ? 1 : 0;
}

// Ordered compaction of the surviving key points of a frame, with the mask
// filter of RemoveInvalidKeyPoints (reference brisk-feature-detector.cc:49-66).
__global__ void __launch_bounds__(256)
compact_kernel(PyramidGeom g, DetectWorkspace ws, const uint8_t* __restrict__ masks, long long mask_frame_stride,
               int mask_pitch, KeyPoint* __restrict__ out, int* __restrict__ counts, int kp_cap) {
  __shared__ int s_warp[8];
  __shared__ int s_base;
  const int frame = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = min(ws.layer_start[(long long)frame * (kMaxLayers + 1) + g.n_layers], ws.corner_cap);
  const KeyPoint* src = ws.kp_tmp + (long long)frame * ws.corner_cap;
  const uint8_t* valid = ws.kp_valid + (long long)frame * ws.corner_cap;
  const uint8_t* mask = masks ? masks + (long long)frame * mask_frame_stride : nullptr;
  KeyPoint* dst = out + (long long)frame * kp_cap;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 256) {
    const int k = base + tid;
    bool keep = k < n && valid[k];
    KeyPoint kp;
    if (keep) {
      kp = src[k];
      if (mask && mask[(long long)(int)(kp.y + 0.5f) * mask_pitch + (int)(kp.x + 0.5f)] == 0) keep = false;
    }
    const uint32_t m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    int before = s_base;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    const int pos = before + __popc(m & ((1u << lane) - 1));
    if (keep && pos < kp_cap) dst[pos] = kp;
    __syncthreads();
    if (tid == 0) { int t = 0; for (int w = 0; w < 8; ++w) t += s_warp[w]; s_base += t; }
    __syncthreads();
  }
  if (tid == 0) counts[frame] = s_base;  // may exceed kp_cap: caller reports truncation
}

// Debug: dense FAST 9-16 / 5-8 score planes of layer 0 (threshold 1, border rules
// of brisk-layer.cc:118-145), for parity tests of the closed-form score.
__global__ void __launch_bounds__(256)
dense_scores_kernel(LayerGeom L, const uint8_t* __restrict__ img, uint8_t* __restrict__ out916, uint8_t* __restrict__ out58) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= L.w || y >= L.h) return;
  const LayerView v{img, nullptr, nullptr, L.w, L.h, L.pitch, L.scale, L.offset};
  out916[(long long)y * L.w + x] = in_border(v, x, y) ? 0 : (uint8_t)fastF(v, x, y);
  out58[(long long)y * L.w + x] = (uint8_t)score58(v, x, y);
}

cudaError_t launch_dense_scores(const LayerGeom& L, const uint8_t* img, uint8_t* out916, uint8_t* out58, cudaStream_t stream) {
  dim3 grid((L.w + 255) / 256, L.h);
  dense_scores_kernel<<<grid, 256, 0, stream>>>(L, img, out916, out58);
  return cudaGetLastError();
}

cudaError_t launch_agast_nms(const PyramidGeom& g, const DetectWorkspace& ws, int n_frames, const uint8_t* masks,
                             long long mask_frame_stride, int mask_pitch, KeyPoint* out, int* counts, int kp_cap,
                             int* error_flag, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(ws.bm, 0, (size_t)n_frames * g.frame_elems, stream);
  if (e != cudaSuccess) return e;
  dim3 grid((ws.corner_cap + 127) / 128, n_frames);
  nms_prefix_kernel<<<grid, 128, 0, stream>>>(g, ws);
  nms_checks_kernel<<<grid, 128, 0, stream>>>(g, ws);
  nms_chain_kernel<<<n_frames, kChainThreads, 0, stream>>>(g, ws, error_flag);
  refine_kernel<<<grid, 128, 0, stream>>>(g, ws);
  compact_kernel<<<n_frames, 256, 0, stream>>>(g, ws, masks, mask_frame_stride, mask_pitch, out, counts, kp_cap);
  return cudaGetLastError();
}

}  // namespace briskb200
