// Scale-space NMS and refinement kernels of the AGAST path (sm_100a); the
// per-corner logic lives in nms_logic.cuh (shared with the host-side tests).
//
// Replaces BriskScaleSpace::GetKeypoints and friends (reference
// brisk/src/brisk-scale-space.cc:92-1364).  Launch sequence per batch:
//   nms_prefix_kernel  one thread per raw corner   IsMax2D's 8 comparisons; flags tying corners
//   nms_checks_kernel  one thread per raw corner   above / below scale checks (pure) and, for
//                                                  corners accepted without a tie, the footprint
//                                                  they leave on the layer above
//   nms_chain_kernel   one CTA per frame           layer by layer: the tying corners only, one
//                                                  warp per corner in raster order, each waiting
//                                                  for its raster-earlier tying neighbours
//   refine_kernel      one thread per raw corner   sub-pixel / scale refinement
//   compact_kernel     one CTA per frame           ordered compaction (+ mask filter)
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

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

__device__ __forceinline__ LayerView make_view(const PyramidGeom& g, const DetectWorkspace& ws, int frame, int layer) {
  const long long fo = (long long)frame * g.frame_elems;
  const LayerGeom& L = g.L[layer];
  return LayerView{ws.pyr + fo + L.off, ws.cm + fo + L.off, ws.bm + fo + L.off, L.w, L.h, L.pitch, L.scale, L.offset};
}

__device__ __forceinline__ void unpack_corner(uint32_t c, int* x, int* y, int* layer) {
  *x = c & 0x1fff; *y = (c >> 13) & 0x1fff; *layer = c >> 26;
}

#ifndef BRISK_PREFIX_MINB
#define BRISK_PREFIX_MINB 5
#endif
#ifndef BRISK_CHECKS_MINB
#define BRISK_CHECKS_MINB 5
#endif
__global__ void __launch_bounds__(128, BRISK_PREFIX_MINB)
nms_prefix_kernel(PyramidGeom g, DetectWorkspace ws) {
  const int frame = blockIdx.y, lane = threadIdx.x & 31;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = min(ws.layer_start[(long long)frame * (kMaxLayers + 1) + g.n_layers], ws.corner_cap);
  if (blockIdx.x * blockDim.x >= n) return;
  bool on = false, tie = false;
  int x = 0, y = 0, layer = 0;
  OwnWindowRegs w;
  w.r[0] = w.r[1] = w.r[2] = w.r[3] = w.r[4] = 0;
  if (k < n) {
    unpack_corner(ws.corners[(long long)frame * ws.corner_cap + k], &x, &y, &layer);
    const uint16_t ev = nms_prefix_mid(make_view(g, ws, frame, layer), x, y, &w);
    on = ev & (kCmAccept | kCmTie);
    tie = ev & kCmTie;
  }
  // the outer rows (-2 and +2) of the tying corners' windows, which only the tie path reads: two row evaluations per
  // tying corner, dealt out over the lanes of the warp (item 2 * i + s = row (s ? +2 : -2) of the warp's i-th tying
  // corner), so that a warp with m ties issues the evaluator's code ceil(2 m / 32) times instead of twice
  {
    const unsigned ties = __ballot_sync(0xffffffffu, tie);
    const int n_items = 2 * __popc(ties);
    const int my_first = 2 * __popc(ties & ((1u << lane) - 1u));   // item index of this lane's row -2 (if it ties)
    for (int base = 0; base < n_items; base += 32) {
      const int item = base + lane;
      const bool active = item < n_items;
      const int src = active ? (int)__fns(ties, 0, (item >> 1) + 1) : 0;
      const int sx = __shfl_sync(0xffffffffu, x, src), sy = __shfl_sync(0xffffffffu, y, src), sl = __shfl_sync(0xffffffffu, layer, src);
      unsigned long long rowv = 0;
      if (active) {
        OwnWindowRegs t;
        t.r[0] = t.r[4] = 0;
        const int wy = (item & 1) ? 2 : -2;
        own_window_rows(make_view(g, ws, frame, sl), sx, sy, wy, wy, &t);
        rowv = (item & 1) ? t.r[4] : t.r[0];
      }
      // hand the rows back to their owners
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int want = my_first + s - base;
        const bool mine = tie && want >= 0 && want < 32;
        const unsigned lo = __shfl_sync(0xffffffffu, (unsigned)rowv, mine ? want : 0);
        const unsigned hi = __shfl_sync(0xffffffffu, (unsigned)(rowv >> 32), mine ? want : 0);
        if (mine) w.r[s ? 4 : 0] = ((unsigned long long)hi << 32) | lo;
      }
    }
  }
  if (on) {
    // the score window: the chain kernel reads all of it (tying corners), refine_kernel its inner 3x3
    uint32_t wds[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 25; ++i) wds[i >> 2] |= (uint32_t)((w.r[i / 5] >> (8 * (i % 5))) & 0xffull) << (8 * (i & 3));
    uint4* dst = reinterpret_cast<uint4*>(ws.fwin + ((long long)frame * ws.corner_cap + k) * 32);
    dst[0] = make_uint4(wds[0], wds[1], wds[2], wds[3]);
    dst[1] = make_uint4(wds[4], wds[5], wds[6], wds[7]);
  }
  // per-layer list of the tying corners (any order) for the chain kernel, in the frame's key-point scratch, which
  // is free until refine_kernel runs; layer l's list starts at its first corner slot.  One atomic per warp and layer
  // (the corners of a warp are consecutive in the layer-major list: nearly always one layer).
  unsigned todo = __ballot_sync(0xffffffffu, tie);
  while (todo) {
    const int ll = __shfl_sync(0xffffffffu, layer, __ffs(todo) - 1);
    const unsigned peers = __ballot_sync(0xffffffffu, tie && layer == ll);
    int base = 0;
    if (lane == __ffs(peers) - 1) base = atomicAdd(&ws.n_ties[frame * kTieStride + ll], __popc(peers));
    base = __shfl_sync(0xffffffffu, base, __ffs(peers) - 1);
    if (tie && layer == ll)
      reinterpret_cast<int2*>(ws.kp_tmp + (long long)frame * ws.corner_cap)[ws.layer_start[(long long)frame * (kMaxLayers + 1) + ll] + base +
                                                                             __popc(peers & ((1u << lane) - 1u))] = make_int2(k, x | (y << 16));
    todo &= ~peers;
  }
  // corners that survive (accepted or tying) go on to the scale checks: compacted (order within a warp
  // kept, warps in any order), so that nms_checks_kernel runs full warps
  const unsigned bal = __ballot_sync(0xffffffffu, on);
  if (bal) {
    const int leader = __ffs(bal) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(&ws.n_ties[frame * kTieStride + kMaxLayers], __popc(bal));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (on) ws.surv[(long long)frame * ws.corner_cap + base + __popc(bal & ((1u << lane) - 1u))] = k;
  }
}

// Warp-cooperative IsMax2D tie path for one tying corner (same result as nms_tie_decide, which
// states the rule pixel by pixel).  The 64 corner-map entries of the 8x8 window are staged by the
// lanes (two each); lanes 0..24 own one pixel of the 5x5 neighbourhood each and reconstruct its
// cache state together, stepping through the raster-earlier corners of the window (a handful,
// found with two ballots) -- under both assumptions about the pending verdicts.  The 3x3 binomial
// sums around the centre and around every tying neighbour are formed with shuffles, as intervals.
// Returns 1 accept, 0 reject, -1 not decidable yet; uniform across the warp.
// Everything warp_tie_decide reads from global memory for one corner, per lane: two entries of the
// 8x8 corner-map window, the FAST score and the touch mark of the lane's pixel of the 5x5
// neighbourhood, and the scan footprint for warp_mark_above.  Loaded one corner ahead of its use.
struct TieLoads {
  uint16_t we[2];
  unsigned F;  // (lanes >= 25: a byte of the record's padding, never used)
  int bmv;
  int2 fp;  // CheckResult::above_steps, above_argmax
};

// Per-lane offsets (in pixels of the layer's planes) of what a lane fetches for a corner at offset 0: its two entries
// of the 8x8 corner-map window and its pixel of the 5x5 neighbourhood.  They only change with the layer's pitch.
struct TieLaneOffsets { int w0, w1, q; };
__device__ __forceinline__ TieLaneOffsets tie_lane_offsets(int pitch) {
  const int lane = threadIdx.x & 31;
  TieLaneOffsets o;
  o.w0 = ((lane >> 3) - 4) * pitch + (lane & 7) - 4;
  o.w1 = o.w0 + 4 * pitch;                       // entry lane + 32: four rows further down
  o.q = (lane / 5 - 2) * pitch + lane % 5 - 2;   // (lanes 0..24)
  return o;
}

__device__ __forceinline__ void tie_issue_loads(const LayerView& L, const TieLaneOffsets& lo, int mode, int x, int y,
                                                const uint8_t* __restrict__ fwin, const float* __restrict__ checks, TieLoads* t) {
  const int lane = threadIdx.x & 31;
  const int base = y * L.pitch + x;   // (a layer plane is far below 2^31 pixels)
  // (every lane loads a byte of the corner's 32-byte window record, so that no conversion or merge instruction has to
  // wait for the load right here: the value is used one corner later)
  t->F = __ldg(fwin + lane);
  t->fp = mode == kModeMid ? *reinterpret_cast<const int2*>(checks + 6) : make_int2(0, 0);
  if (x >= 7 && y >= 7 && x < L.w - 7 && y < L.h - 7) {
    // interior corner (uniform across the warp): the whole 8x8 window and the 5x5 neighbourhood lie inside the border
    t->we[0] = L.cm[base + lo.w0];
    t->we[1] = L.cm[base + lo.w1];
    t->bmv = lane < 25 ? L.bm[base + lo.q] : 0;
    return;
  }
  const int ox = lane % 5 - 2, oy = lane / 5 - 2;  // lanes 0..24 <-> 5x5 offsets, row-major
  t->we[0] = 0; t->we[1] = 0; t->bmv = 0;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int i = lane + 32 * h;
    const int px = x - 4 + (i & 7), py = y - 4 + (i >> 3);
    if (py >= 3 && px >= 3 && px < L.w - 3 && py < L.h - 3) t->we[h] = L.cm[(long long)py * L.pitch + px];
  }
  if (lane < 25 && !in_border(L, x + ox, y + oy)) t->bmv = L.bm[(long long)(y + oy) * L.pitch + x + ox];
}

template <int mode>
__device__ __forceinline__ int warp_tie_decide(const LayerView& L, int x, int y, const TieLoads& ld,
                                               uint16_t* s_win /* 64 entries, this warp's */) {
  constexpr unsigned kFull = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int ox = lane % 5 - 2, oy = lane / 5 - 2;  // lanes 0..24 <-> 5x5 offsets, row-major
  const int F = lane < 25 ? (int)ld.F : 0;
  s_win[lane] = ld.we[0];
  s_win[lane + 32] = ld.we[1];
  // raster-earlier corners of the window (entry 36 is the corner itself): the only entries cache_state looks at
  const bool c0 = ld.we[0] & kCmT, c1 = (ld.we[1] & kCmT) && lane < 4;
  unsigned long long earlier = (unsigned long long)__ballot_sync(kFull, c0) | ((unsigned long long)__ballot_sync(kFull, c1) << 32);
  const bool pending = __any_sync(kFull, (c0 && !(ld.we[0] & kCmDecided)) || (c1 && !(ld.we[1] & kCmDecided)));
  __syncwarp();
  const int center = s_win[4 * 8 + 4] & kCmT;
  const bool ring1 = ox >= -1 && ox <= 1 && oy >= -1 && oy <= 1;
  const int tq = lane < 25 ? (s_win[(oy + 4) * 8 + ox + 4] & kCmT) : 0;  // 0 on border pixels
  // cache_state of the lane's pixel q, for all lanes at once: one step per raster-earlier corner p of the
  // window (uniform), in raster order; pending verdicts taken as reject (0) / accept (1)
  int sticky0 = ld.bmv != 0 ? 1 : 0, sticky1 = sticky0;  // (ints: bools get packed into bytes of one register)
  int last0 = sticky0, last1 = last0;
  while (earlier) {
    const int i = __ffsll((long long)earlier) - 1;
    earlier &= earlier - 1;
    const int e = s_win[i];
    const int t = e & kCmT, calls = (e & kCmCalls) >> kCmCallsShift;
    const bool acc0 = e & kCmAccept, acc1 = acc0 || !(e & kCmDecided), chk = e & kCmChecks;
    const int pox = ox + 4 - (i & 7), poy = oy + 4 - (i >> 3);  // q - p; p can touch q when both are in [-1,2]
    const bool inblk = (unsigned)(pox + 1) <= 3u && (unsigned)(poy + 1) <= 3u;
    const bool near = inblk && pox <= 1 && poy <= 1;
    // IsMax2D neighbour look-up of p with threshold T(p), if p got that far (isMax2dIndex as a nibble table)
    const int j = (int)((0x534150627ull >> (4 * (((poy + 1) * 3 + pox + 1) & 15))) & 0xfull);
    const bool look = near && j < calls;
    bool patch;  // q inside the patch an accepted p looks up with threshold 1
    if (mode == kModeSingle) patch = inblk;
    else if (mode == kModeLast) patch = inblk && ((pox >= 0 && poy >= 0 && near) || chk);
    else patch = near && chk;
    if (look) { last0 = t; last1 = t; if (t <= F) { sticky0 = 1; sticky1 = 1; } }
    if (patch && acc0) { sticky0 = 1; last0 = 1; }
    if (patch && acc1) { sticky1 = 1; last1 = 1; }
  }
  int st0 = 0, st1 = 0;
  if (lane < 25 && !tq && F >= 1 && !in_border(L, x + ox, y + oy)) {
    st0 = F > 2 ? (sticky0 ? F : 0) : ((last0 != 0 && last0 <= F) ? F : 0);
    st1 = F > 2 ? (sticky1 ? F : 0) : ((last1 != 0 && last1 <= F) ? F : 0);
  }
  // what the tie path sees (tie_pixel_value): own score, look-up results on the 8 neighbours, raw cache bytes outside
  int v, v1;
  if (lane == 12) v = v1 = center;
  else if (tq) v = v1 = ring1 ? F : tq;  // a neighbouring corner: its T (stored in fwin by nms_prefix)
  else if (ring1) { v = st0 > 2 ? st0 : (F >= center ? F : 0); v1 = st1 > 2 ? st1 : (F >= center ? F : 0); }
  else { v = st0; v1 = st1; }
  // binomial 3x3 sums centred on every lane's own pixel (meaningful for the inner 3x3 lanes), as intervals:
  // `sum` with the pending verdicts taken as reject, `sum1` as accept (nms_tie_decide)
  int sum = 0, sum1 = 0;
#pragma unroll
  for (int wy = -1; wy <= 1; ++wy)
#pragma unroll
    for (int wx = -1; wx <= 1; ++wx) {
      const int src = lane + wy * 5 + wx;
      sum += ((wx == 0 ? 2 : 1) * (wy == 0 ? 2 : 1)) * __shfl_sync(kFull, v, src & 31);
    }
  if (pending) {
#pragma unroll
    for (int wy = -1; wy <= 1; ++wy)
#pragma unroll
      for (int wx = -1; wx <= 1; ++wx) {
        const int src = lane + wy * 5 + wx;
        sum1 += ((wx == 0 ? 2 : 1) * (wy == 0 ? 2 : 1)) * __shfl_sync(kFull, v1, src & 31);
      }
  } else {
    sum1 = sum;
  }
  const int s_lo = __shfl_sync(kFull, sum, 12), s_hi = __shfl_sync(kFull, sum1, 12);
  const bool inner = lane < 25 && ring1 && lane != 12;
  // a certainly tying neighbour that certainly beats the centre rejects; a possibly tying one that possibly
  // does leaves the verdict open
  const bool lost = inner && v == center && v1 == center && sum > s_hi;
  const bool may_lose = inner && (v == center || v1 == center) && sum1 > s_lo;
  const int verdict = __any_sync(kFull, lost) ? 0 : (__any_sync(kFull, may_lose) ? -1 : 1);
  __syncwarp();
  return verdict;
}

// mark_above1 as a strided loop: the scan visits a (columns x rows) grid of positions in row-major order --
// columns x_1, the integers in (x_1, x1], x1; rows alike -- and `steps` of them were evaluated; position n is
// replayed by whoever owns n = first, first + stride, ...  (first = lane, stride = 32: one corner spread over a warp,
// the footprint being uniform across it; first = 0, stride = 1: one thread on its own corner.)
__device__ __forceinline__ void mark_above_strided(const LayerView& nb, int layer, int x, int y, int above_steps, int above_argmax,
                                                   int first, int stride) {
  float x_1, x1, y_1, y1;
  above_patch(layer, x, y, &x_1, &x1, &y_1, &y1);
  const int xb = (int)(x_1 + 1), xe = (int)x1, yb = (int)(y_1 + 1), ye = (int)y1;
  const int ncols = imax(xe - xb + 1, 0) + 2, nrows = imax(ye - yb + 1, 0) + 2;
  const int steps = above_steps & 0xff;
#pragma unroll 1
  for (int n = first; n < steps; n += stride) {
    const int c = n % ncols, rr = n / ncols;
    const bool xi = c > 0 && c < ncols - 1, yi = rr > 0 && rr < nrows - 1;
    const float xf = c == 0 ? x_1 : (xi ? (float)(xb + c - 1) : x1);
    const float yf = rr == 0 ? y_1 : (yi ? (float)(yb + rr - 1) : y1);
    if (xi && yi) mark_px(nb, xb + c - 1, yb + rr - 1);
    else mark_cell(nb, xf, yf);
  }
  if ((above_steps >> 8) & 1) {
#pragma unroll 1
    for (int q = first; q < 9; q += stride) mark_px(nb, (above_argmax & 0xffff) + q % 3 - 1, (above_argmax >> 16) + q / 3 - 1);
  }
}
__device__ __forceinline__ void warp_mark_above(const LayerView& nb, int layer, int x, int y, int above_steps, int above_argmax) {
  mark_above_strided(nb, layer, x, y, above_steps, above_argmax, threadIdx.x & 31, 32);
}

__global__ void __launch_bounds__(128, BRISK_CHECKS_MINB)
nms_checks_kernel(PyramidGeom g, DetectWorkspace ws) {
  const int frame = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = ws.n_ties[frame * kTieStride + kMaxLayers];  // corners that passed IsMax2D's comparisons
  if (blockIdx.x * blockDim.x >= n) return;
  int x = 0, y = 0, layer = 0;
  CheckResult r;
  bool mark = false;
  if (i < n) {
    const int k = ws.surv[(long long)frame * ws.corner_cap + i];
    unpack_corner(ws.corners[(long long)frame * ws.corner_cap + k], &x, &y, &layer);
    const LayerView own = make_view(g, ws, frame, layer);
    uint16_t* e = own.cm + (long long)y * own.pitch + x;
    const uint16_t ev = *e;
    // only the corner's own layer and its two neighbours are looked at
    const LayerView below = make_view(g, ws, frame, layer > 0 ? layer - 1 : 0);
    const LayerView above = make_view(g, ws, frame, layer + 1 < g.n_layers ? layer + 1 : layer);
    const bool ok = nms_checks3(below, own, above, g.n_layers, layer, x, y, &r);
    if (ok) *e = ev | kCmChecks;
    // kept even when the checks fail: the footprint of the scan of the layer above is needed by the chain kernel
    *reinterpret_cast<CheckResult*>(ws.checks + ((long long)frame * ws.corner_cap + k) * 8) = r;
    // A corner accepted without a tie leaves its footprint on the layer above right away (the touch map is
    // only read by the chain kernel); tying corners do so once they are resolved.
    mark = (ev & kCmAccept) && g.n_layers > 1 && layer < g.n_layers - 1;
  }
  // every thread replays its own corner's scan (a few to 25 byte stores): a warp issues the longest of its 32
  // replays once, where going through the corners one at a time with one lane per position issued all of them
  if (mark) mark_above_strided(make_view(g, ws, frame, layer + 1), layer, x, y, r.above_steps, r.above_argmax, 0, 1);
}

// One CTA per frame, layers in order (layer i+1 needs the touch marks that layer i's accepted
// corners leave on it).  Only the tying corners are handled here (nms_prefix_kernel listed them per
// layer).  A layer is resolved in passes, one warp per corner: a corner whose verdict does not depend on
// a pending one is decided and published (one 16-bit store; a concurrent reader may or may not see it,
// both are fine), the others form the list of the next pass.  The raster-earliest pending corner is
// always decidable, so every pass makes progress.
// Two 512-thread CTAs per SM: the passes of one frame (barriers, short lists on the upper layers) leave the SM idle
// part of the time; a second frame fills those gaps (12.8 -> 9.0 ms per 1024 frames against one 1024-thread CTA).
#ifndef BRISK_CHAIN_THREADS
#define BRISK_CHAIN_THREADS 512
#endif
constexpr int kChainThreads = BRISK_CHAIN_THREADS;
constexpr int kChainWarps = kChainThreads / 32;
__global__ void __launch_bounds__(kChainThreads, 1024 / kChainThreads)
nms_chain_kernel(PyramidGeom g, DetectWorkspace ws, int* __restrict__ error_flag) {
  __shared__ FrameViews fv;
  __shared__ int s_next[2];
  __shared__ uint16_t s_win[kChainWarps][64];
  const int frame = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int* ls = ws.layer_start + (long long)frame * (kMaxLayers + 1);
  if (ls[g.n_layers] > ws.corner_cap) return;  // corner list truncated: the call fails with BRISK_ERR_CAPACITY anyway
  if (tid == 0) { make_views(g, ws, frame, &fv); s_next[0] = 0; s_next[1] = 0; }
  __syncthreads();
  // the key-point scratch of the frame is free until refine_kernel runs: it holds the tie lists (corner
  // slot, x | y << 16; layer l's list starts at its first corner slot) and the lists of the next passes
  int2* lists = reinterpret_cast<int2*>(ws.kp_tmp + (long long)frame * ws.corner_cap);
  const long long fslot = (long long)frame * ws.corner_cap;
  for (int layer = 0; layer < g.n_layers; ++layer) {
    const int mode = g.n_layers == 1 ? kModeSingle : (layer == g.n_layers - 1 ? kModeLast : kModeMid);
    const LayerView& L = fv.v[layer];
    const TieLaneOffsets lane_off = tie_lane_offsets(L.pitch);
    int n = ws.n_ties[frame * kTieStride + layer];
    const int2* cur = lists + ls[layer];
    for (int pass = 0; n > 0; ++pass) {
      int2* nxt = lists + (1 + (pass & 1)) * (long long)ws.corner_cap;  // never the list being read
      // software pipeline: the list entry is fetched two corners ahead, the corner's data one ahead
      // (a window read early may miss a verdict of this pass: that can only defer the corner, never
      // change its verdict, and everything decided in earlier passes is seen)
      int2 ent = warp < n ? cur[warp] : make_int2(0, 0);
      int2 ent1 = warp + kChainWarps < n ? cur[warp + kChainWarps] : make_int2(0, 0);
      TieLoads ld;
      if (warp < n) tie_issue_loads(L, lane_off, mode, ent.y & 0xffff, ent.y >> 16, ws.fwin + (fslot + ent.x) * 32, ws.checks + (fslot + ent.x) * 8, &ld);
      for (int i = warp; i < n; i += kChainWarps) {
        TieLoads ld1;
        int2 ent2 = make_int2(0, 0);
        if (i + kChainWarps < n) tie_issue_loads(L, lane_off, mode, ent1.y & 0xffff, ent1.y >> 16, ws.fwin + (fslot + ent1.x) * 32, ws.checks + (fslot + ent1.x) * 8, &ld1);
        if (i + 2 * kChainWarps < n) ent2 = cur[i + 2 * kChainWarps];
        const int x = ent.y & 0xffff, y = ent.y >> 16;
        const int verdict = mode == kModeMid ? warp_tie_decide<kModeMid>(L, x, y, ld, s_win[warp])
                          : mode == kModeLast ? warp_tie_decide<kModeLast>(L, x, y, ld, s_win[warp])
                                              : warp_tie_decide<kModeSingle>(L, x, y, ld, s_win[warp]);
        if (lane == 0) {
          if (verdict < 0) nxt[atomicAdd(&s_next[pass & 1], 1)] = ent;
          else {
            uint16_t* e = L.cm + (long long)y * L.pitch + x;
            *e = s_win[warp][4 * 8 + 4] | (uint16_t)(kCmDecided | (verdict ? kCmAccept : 0));  // the window's centre is the corner's own entry
          }
        }
        // footprint of the accepted corner on the layer above
        if (verdict > 0 && mode == kModeMid) warp_mark_above(fv.v[layer + 1], layer, x, y, ld.fp.x, ld.fp.y);
        __syncwarp();
        ent = ent1; ent1 = ent2; ld = ld1;
      }
      __syncthreads();
      const int left = s_next[pass & 1];
      if (tid == 0) s_next[(pass + 1) & 1] = 0;
      __syncthreads();
      if (left >= n) {  // no progress: cannot happen
        if (tid == 0) { atomicExch(error_flag, 2); s_next[pass & 1] = 0; }
        break;
      }
      n = left;
      cur = nxt;
    }
    __syncthreads();
  }
}

// One thread per corner that passed IsMax2D's comparisons (the compacted list of nms_prefix_kernel); the
// validity bytes of all other slots were cleared by launch_agast_nms.
__global__ void __launch_bounds__(128)
refine_kernel(PyramidGeom g, DetectWorkspace ws) {
  const int frame = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ws.n_ties[frame * kTieStride + kMaxLayers]) return;
  const int k = ws.surv[(long long)frame * ws.corner_cap + i];
  const long long slot = (long long)frame * ws.corner_cap + k;
  int x, y, layer;
  unpack_corner(ws.corners[slot], &x, &y, &layer);
  const LayerView own = make_view(g, ws, frame, layer);
  const uint16_t e = own.cm[(long long)y * own.pitch + x];
  if (!(e & kCmAccept) || !(e & kCmChecks)) return;
  const CheckResult r = *reinterpret_cast<const CheckResult*>(ws.checks + slot * 8);
  KeyPoint kp;
  if (refine_emit1(own, g.n_layers, layer, x, y, r, &kp, ws.fwin + slot * 32)) {
    ws.kp_tmp[slot] = kp;
    ws.kp_valid[slot] = 1;
  }
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

// ---------------------------------------------------------------------------
// BriskFeatureDetector::ComputeScale (reference brisk-feature-detector.cc:87-92): GetKeypoints with
// caller-provided key points (brisk-scale-space.cc:104-124).  One thread per (frame, layer, provided
// point); passes (nms_logic.cuh, "Provided key points"):
//   0  count the points every layer keeps (n_ties[frame][layer]); a layer that keeps none is a FALLBACK
//      layer: the reference runs its detector there (brisk-layer.cc:103-105; lower threshold 0) and
//      treats the corners like provided points -- the caller then runs the detect kernel with lower = 0
//      and the corner lists before going on
//   1  the threshold-0 look-ups        2  the score write of GetAgastPoints
//   3  scale checks + refinement into the frame's key-point scratch; layer l's slots start at
//      base[frame][l] (in_max slots for a layer with provided points, one per corner for a fallback
//      layer), so that compact_kernel packs them in the reference's order (layers, then input order).
// ---------------------------------------------------------------------------
template <int PASS>
__global__ void __launch_bounds__(128)
provided_kernel(PyramidGeom g, DetectWorkspace ws, const KeyPoint* __restrict__ in, const int* __restrict__ in_counts, int in_cap,
                const int* __restrict__ base) {
  const int frame = blockIdx.z, layer = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_in = min(in_counts[frame], in_cap);
  int* kept = ws.n_ties + frame * kTieStride + layer;
  const LayerView L = make_view(g, ws, frame, layer);
  float px = 0.0f, py = 0.0f;
  KeyPoint kin;
  bool inside = false;
  if (j < n_in) {
    kin = in[(long long)frame * in_cap + j];
    inside = provided_to_layer(L, kin.x, kin.y, &px, &py);
  }
  if (PASS == 0) {
    const unsigned m = __ballot_sync(0xffffffffu, inside);
    if (m && (threadIdx.x & 31) == 0) atomicAdd(kept, __popc(m));
    return;
  }
  if (!inside) return;
  if (PASS == 1) provided_touch(L, px, py);
  else if (PASS == 2) provided_stamp(L, px, py);
  else {
    const LayerView below = make_view(g, ws, frame, layer > 0 ? layer - 1 : 0);
    const LayerView above = make_view(g, ws, frame, layer + 1 < g.n_layers ? layer + 1 : layer);
    KeyPoint kp;
    kp.class_id = kin.class_id;
    if (provided_refine(below, L, above, g.n_layers, layer, px, py, &kp)) {
      const long long slot = (long long)frame * ws.corner_cap + base[frame * (kMaxLayers + 1) + layer] + j;
      ws.kp_tmp[slot] = kp;
      ws.kp_valid[slot] = 1;
    }
  }
}

// One thread per frame: first slot of every layer (and the total, which compact_kernel reads).
__global__ void provided_layout_kernel(PyramidGeom g, DetectWorkspace ws, int n_frames, int in_max, int with_fallback,
                                       int* __restrict__ base, int* __restrict__ error_flag) {
  const int frame = blockIdx.x * blockDim.x + threadIdx.x;
  if (frame >= n_frames) return;
  const int* ls = ws.layer_start + (long long)frame * (kMaxLayers + 1);
  int total = 0;
  for (int l = 0; l < g.n_layers; ++l) {
    base[frame * (kMaxLayers + 1) + l] = total;
    if (ws.n_ties[frame * kTieStride + l] > 0) total += in_max;
    else if (with_fallback) total += ls[l + 1] - ls[l];
    else atomicExch(error_flag, 3);
  }
  base[frame * (kMaxLayers + 1) + g.n_layers] = total;
  if (total > ws.corner_cap || (with_fallback && ls[g.n_layers] > ws.corner_cap)) atomicExch(error_flag, 1);
}

// After the detect kernel ran on every layer: layers with provided points start from an empty score cache;
// on fallback layers a corner whose score is <= 2 does not stay cached (brisk-layer.cc:124-126) -- its entry
// goes, its place in the corner list stays.
__global__ void __launch_bounds__(256)
provided_clear_kernel(PyramidGeom g, DetectWorkspace ws) {
  const int frame = blockIdx.z, layer = blockIdx.y;
  const LayerGeom& G = g.L[layer];
  uint16_t* cm = ws.cm + (long long)frame * g.frame_elems + G.off;
  const bool fallback = ws.n_ties[frame * kTieStride + layer] == 0;
  const long long n = (long long)G.pitch * G.h;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const uint16_t v = cm[i];
    if (v && (!fallback || (v & kCmT) <= 2)) cm[i] = 0;
  }
}

// Fallback layers: one thread per detected corner.
__global__ void __launch_bounds__(128)
provided_corner_kernel(PyramidGeom g, DetectWorkspace ws, const int* __restrict__ base) {
  const int frame = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int* ls = ws.layer_start + (long long)frame * (kMaxLayers + 1);
  if (k >= min(ls[g.n_layers], ws.corner_cap)) return;
  int x, y, layer;
  unpack_corner(ws.corners[(long long)frame * ws.corner_cap + k], &x, &y, &layer);
  if (ws.n_ties[frame * kTieStride + layer] > 0) return;
  const LayerView L = make_view(g, ws, frame, layer);
  const LayerView below = make_view(g, ws, frame, layer > 0 ? layer - 1 : 0);
  const LayerView above = make_view(g, ws, frame, layer + 1 < g.n_layers ? layer + 1 : layer);
  KeyPoint kp;
  kp.class_id = -1;
  if (provided_refine(below, L, above, g.n_layers, layer, (float)x, (float)y, &kp)) {
    const long long slot = (long long)frame * ws.corner_cap + base[frame * (kMaxLayers + 1) + layer] + (k - ls[layer]);
    if (slot < (long long)(frame + 1) * ws.corner_cap) {
      ws.kp_tmp[slot] = kp;
      ws.kp_valid[slot] = 1;
    }
  }
}

cudaError_t launch_provided_count(const PyramidGeom& g, const DetectWorkspace& ws, int n_frames, const KeyPoint* in,
                                  const int* in_counts, int in_cap, int in_max, cudaStream_t stream) {
  if (in_max <= 0) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemsetAsync(ws.n_ties, 0, (size_t)n_frames * kTieStride * sizeof(int), stream);
  if (e != cudaSuccess) return e;
  dim3 grid((in_max + 127) / 128, g.n_layers, n_frames);
  provided_kernel<0><<<grid, 128, 0, stream>>>(g, ws, in, in_counts, in_cap, nullptr);
  return cudaGetLastError();
}

cudaError_t launch_provided_scale(const PyramidGeom& g, const DetectWorkspace& ws, int n_frames, const KeyPoint* in,
                                  const int* in_counts, int in_cap, int in_max, int with_fallback, int* base, KeyPoint* out,
                                  int* counts, int kp_cap, int* error_flag, cudaStream_t stream) {
  if (in_max <= 0) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemsetAsync(ws.kp_valid, 0, (size_t)n_frames * ws.corner_cap, stream);
  if (e != cudaSuccess) return e;
  if (with_fallback) {
    dim3 cgrid(64, g.n_layers, n_frames);
    provided_clear_kernel<<<cgrid, 256, 0, stream>>>(g, ws);
  } else {
    e = cudaMemsetAsync(ws.cm, 0, (size_t)n_frames * g.frame_elems * sizeof(uint16_t), stream);
    if (e != cudaSuccess) return e;
  }
  provided_layout_kernel<<<(n_frames + 127) / 128, 128, 0, stream>>>(g, ws, n_frames, in_max, with_fallback, base, error_flag);
  dim3 grid((in_max + 127) / 128, g.n_layers, n_frames);
  provided_kernel<1><<<grid, 128, 0, stream>>>(g, ws, in, in_counts, in_cap, base);
  provided_kernel<2><<<grid, 128, 0, stream>>>(g, ws, in, in_counts, in_cap, base);
  provided_kernel<3><<<grid, 128, 0, stream>>>(g, ws, in, in_counts, in_cap, base);
  if (with_fallback) {
    dim3 kgrid((ws.corner_cap + 127) / 128, n_frames);
    provided_corner_kernel<<<kgrid, 128, 0, stream>>>(g, ws, base);
  }
  DetectWorkspace cw = ws;
  cw.layer_start = base;  // compact_kernel packs the first base[frame][n_layers] slots
  compact_kernel<<<n_frames, 256, 0, stream>>>(g, cw, nullptr, 0, 0, out, counts, kp_cap);
  return cudaGetLastError();
}

// Debug: dense FAST 9-16 / 5-8 score planes of layer 0 (threshold 1, border rules
// of brisk-layer.cc:118-145), for parity tests of the closed-form score.
__global__ void __launch_bounds__(256)
dense_scores_kernel(LayerGeom L, const uint8_t* __restrict__ img, uint8_t* __restrict__ out916, uint8_t* __restrict__ out58) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= L.w || y >= L.h) return;
  const LayerView v{img, nullptr, nullptr, L.w, L.h, L.pitch, L.scale, L.offset};
  // the packed row evaluator of the NMS kernels (fast_packed.cuh), at every alignment: pixel x is lane (x % 6) of the run
  // starting at x - x % 6 for odd rows, and lane (x % 4) of a four-pixel run for even rows
  uint32_t f[3];
  int lane;
  if (y & 1) { lane = x % 6; fast916_row<3>(img, L.pitch, L.h, x - lane, y, f); }
  else { lane = x % 4; fast916_row<2>(img, L.pitch, L.h, x - lane, y, f); }
  const int packed = (int)((f[lane >> 1] >> ((lane & 1) * 16)) & 0xffffu);
  out916[(long long)y * L.w + x] = in_border(v, x, y) ? 0 : (uint8_t)packed;
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
  e = cudaMemsetAsync(ws.n_ties, 0, (size_t)n_frames * kTieStride * sizeof(int), stream);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(ws.kp_valid, 0, (size_t)n_frames * ws.corner_cap, stream);
  if (e != cudaSuccess) return e;
  dim3 grid((ws.corner_cap + 127) / 128, n_frames);
  // BRISK_B200_NMS_TIMING=1: per-kernel CUDA-event times of this launch sequence on stderr (debug aid; synchronises)
  static const bool timing = getenv("BRISK_B200_NMS_TIMING") != nullptr;
  cudaEvent_t ev[6];
  if (timing) { for (auto& v : ev) cudaEventCreate(&v); cudaEventRecord(ev[0], stream); }
  nms_prefix_kernel<<<grid, 128, 0, stream>>>(g, ws);
  if (timing) cudaEventRecord(ev[1], stream);
  nms_checks_kernel<<<grid, 128, 0, stream>>>(g, ws);
  if (timing) cudaEventRecord(ev[2], stream);
  nms_chain_kernel<<<n_frames, kChainThreads, 0, stream>>>(g, ws, error_flag);
  if (timing) cudaEventRecord(ev[3], stream);
  refine_kernel<<<grid, 128, 0, stream>>>(g, ws);
  if (timing) cudaEventRecord(ev[4], stream);
  compact_kernel<<<n_frames, 256, 0, stream>>>(g, ws, masks, mask_frame_stride, mask_pitch, out, counts, kp_cap);
  if (timing) {
    cudaEventRecord(ev[5], stream);
    cudaStreamSynchronize(stream);
    float ms[5];
    for (int i = 0; i < 5; ++i) cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]);
    fprintf(stderr, "[nms %d frames] prefix %.3f checks %.3f chain %.3f refine %.3f compact %.3f ms\n", n_frames, ms[0], ms[1], ms[2], ms[3], ms[4]);
    for (auto& v : ev) cudaEventDestroy(v);
  }
  return cudaGetLastError();
}

}  // namespace briskb200
