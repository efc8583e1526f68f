// Harris scale-space detector kernels (sm_100a).
//
// Replaces ScaleSpaceFeatureDetector<HarrisScoreCalculator>::detectImpl and its layers
// (reference brisk/include/brisk/scale-space-feature-detector.h:100-128,
// internal/scale-space-layer-inl.h:60-428, brisk/src/harris-scores.cc:53-279,
// brisk/src/harris-score-calculator.cc:57-106, internal/uniformity-enforcement-inl.h:44-194).
//   harris_scores_kernel      tile + halo 2 in shared memory: Scharr products (pmulhw semantics),
//                             3x3 binomial smoothing, integer response -> i32 score plane
//   harris_maxima_*           8-neighbour maxima >= threshold, raster-ordered lists (warp per row)
//   harris_nms3d_kernel       bilinear double-precision reads of the layers above / below
//   harris_sort_kernel        per (frame, layer): compaction + replay of libstdc++'s introsort
//                             (the tie order of equal scores is part of the reference's result)
//   harris_uniformity_kernel  per (frame, layer): greedy occupancy stamping, 31x31 saturating adds
//   harris_emit_kernel        sub-pixel refinement, key points in layer-major order
#include <cuda_runtime.h>

#include "harris_logic.cuh"
#include "kernels.h"

namespace briskb200 {

constexpr int kHsTW = 64, kHsTH = 16, kHsThreads = 256;

__global__ void __launch_bounds__(kHsThreads)
harris_scores_kernel(LayerGeom L, long long frame_elems, const uint8_t* __restrict__ pyr, int* __restrict__ scores) {
  __shared__ uint8_t s_img[kHsTH + 4][kHsTW + 4];
  __shared__ short s_xx[kHsTH + 2][kHsTW + 2], s_yy[kHsTH + 2][kHsTW + 2], s_xy[kHsTH + 2][kHsTW + 2];
  const int x0 = blockIdx.x * kHsTW, y0 = blockIdx.y * kHsTH, frame = blockIdx.z, tid = threadIdx.x;
  const uint8_t* img = pyr + (long long)frame * frame_elems + L.off;
  int* out = scores + (long long)frame * frame_elems + L.off;
  for (int i = tid; i < (kHsTH + 4) * (kHsTW + 4); i += kHsThreads) {
    const int r = i / (kHsTW + 4), c = i - r * (kHsTW + 4);
    const int y = y0 - 2 + r, x = x0 - 2 + c;
    s_img[r][c] = (y >= 0 && y < L.h && x >= 0 && x < L.w) ? img[(long long)y * L.pitch + x] : 0;
  }
  __syncthreads();
  for (int i = tid; i < (kHsTH + 2) * (kHsTW + 2); i += kHsThreads) {
    const int r = i / (kHsTW + 2), c = i - r * (kHsTW + 2);  // product plane (r, c) <-> pixel (y0-1+r, x0-1+c) <-> s_img[r+1][c+1]
    int p[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) p[k] = s_img[r + k / 3][c + k % 3];
    int xx, yy, xy;
    harris_products(p, &xx, &yy, &xy);
    s_xx[r][c] = (short)xx; s_yy[r][c] = (short)yy; s_xy[r][c] = (short)xy;
  }
  __syncthreads();
  for (int i = tid; i < kHsTH * kHsTW; i += kHsThreads) {
    const int r = i / kHsTW, c = i - r * kHsTW;
    const int y = y0 + r, x = x0 + c;
    if (x >= L.w || y >= L.h) continue;
    int v = 0;
    if (x >= 2 && x < L.w - 2 && y >= 2 && y < L.h - 2) {
      int qa[9], qb[9], qc[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) { qa[k] = s_xx[r + k / 3][c + k % 3]; qb[k] = s_yy[r + k / 3][c + k % 3]; qc[k] = s_xy[r + k / 3][c + k % 3]; }
      v = harris_response(harris_smooth(qa), harris_smooth(qb), harris_smooth(qc));
    }
    out[(long long)y * L.pitch + x] = v;
  }
}

__device__ __forceinline__ bool harris_is_max(const int* __restrict__ p, int pitch, int thr) {
  const int c = p[0];
  if (c < thr) return false;
  return !(p[1] > c || p[-1] > c || p[pitch] > c || p[-pitch] > c || p[pitch + 1] > c || p[pitch - 1] > c || p[-pitch + 1] > c || p[-pitch - 1] > c);
}

// warp per row; FILL = false: count into rowcnt, FILL = true: emit at the row's slot
template <bool FILL>
__global__ void __launch_bounds__(256)
harris_maxima_kernel(LayerGeom L, long long frame_elems, const int* __restrict__ scores, int thr, int* __restrict__ rowcnt,
                     int total_rows, int row_off, HPoint* __restrict__ pts, int cap) {
  const int frame = blockIdx.y, lane = threadIdx.x & 31;
  const int y = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (y >= L.h) return;
  int* rc = rowcnt + (long long)frame * total_rows + row_off + y;
  if (y < 2 || y >= L.h - 2) { if (!FILL && lane == 0) *rc = 0; return; }
  const int* row = scores + (long long)frame * frame_elems + L.off + (long long)y * L.pitch;
  int slot = FILL ? *rc : 0;
  HPoint* out = pts + (long long)frame * cap;
  for (int xb = 0; xb < L.w; xb += 32) {
    const int x = xb + lane;
    const bool m = x >= 2 && x < L.w - 2 && harris_is_max(row + x, L.pitch, thr);
    const uint32_t bal = __ballot_sync(0xffffffffu, m);
    if (FILL && m) {
      const int s = slot + __popc(bal & ((1u << lane) - 1));
      if (s < cap) out[s] = HPoint{row[x], (unsigned short)x, (unsigned short)y};
    }
    slot += __popc(bal);
  }
  if (!FILL && lane == 0) *rc = slot;
}

// layer of maximum k of a frame (layer_start is ascending)
__device__ __forceinline__ int layer_of(const int* ls, int n_layers, int k) {
  int l = 0;
  while (l + 1 < n_layers && k >= ls[l + 1]) ++l;
  return l;
}

struct HarrisLayerParams {
  double scale[kMaxLayers], offset[kMaxLayers], scale_above[kMaxLayers], offset_above[kMaxLayers], scale_below[kMaxLayers], offset_below[kMaxLayers];
};

// 3-D non-maximum suppression (scale-space-layer-inl.h:218-366): reject a maximum that is smaller
// than any of 9 bilinear reads of the layer above, or than the read of the layer below (the
// reference's nine "below" reads collapse to one because int(1/scale_below) == 0).
__global__ void __launch_bounds__(128)
harris_nms3d_kernel(PyramidGeom g, HarrisLayerParams hp, const int* __restrict__ scores, const int* __restrict__ layer_start,
                    const HPoint* __restrict__ pts, uint8_t* __restrict__ keep, int cap, int abs_thr) {
  const int frame = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int* ls = layer_start + (long long)frame * (kMaxLayers + 1);
  const int n = min(ls[g.n_layers], cap);
  if (k >= n) return;
  const int layer = layer_of(ls, g.n_layers, k);
  const HPoint p = pts[(long long)frame * cap + k];
  const long long fo = (long long)frame * g.frame_elems;
  bool ok = p.score >= abs_thr;
  if (ok && layer + 1 < g.n_layers) {
    const LayerGeom& A = g.L[layer + 1];
    const int* sa = scores + fo + A.off;
    const double sc = hp.scale_above[layer], of = hp.offset_above[layer];
    const int ox[9] = {0, 1, -1, 0, 0, 1, 1, -1, -1}, oy[9] = {0, 0, 0, 1, -1, 1, -1, 1, -1};
#pragma unroll 1
    for (int j = 0; j < 9 && ok; ++j) {
      const double u = (double)((int)p.x + ox[j]), v = (double)((int)p.y + oy[j]);
      if ((double)p.score < harris_score_bilinear(sa, A.pitch, A.w, A.h, sc * (u + of), sc * (v + of))) ok = false;
    }
  }
  if (ok && layer > 0) {
    const LayerGeom& B = g.L[layer - 1];
    const double sc = hp.scale_below[layer], of = hp.offset_below[layer];
    if ((double)p.score < harris_score_bilinear(scores + fo + B.off, B.pitch, B.w, B.h, sc * ((double)(int)p.x + of), sc * ((double)(int)p.y + of))) ok = false;
  }
  keep[(long long)frame * cap + k] = ok ? 1 : 0;
}

// One CTA per (frame, layer): stable compaction of the kept maxima, then the order std::sort gives them
// (descending score; the arrangement of equal scores is part of the reference's result, so the CTA
// replays libstdc++'s introsort).  Large ranges are partitioned by the whole CTA (cta_partition), the rest
// level by level with one thread per open range; ranges of at most 16 elements get their share of the
// final insertion sort from the thread that produced them.
constexpr int kSortThreads = 256;
constexpr int kSortRanges = 1792;   // open ranges per level held in shared memory; overflow is sorted by its producer
constexpr int kCoopMin = 768;       // ranges above this size are partitioned by the whole CTA
constexpr int kCoopStack = 96;      // > 2 x depth limit of a 2^20-element sort

struct SortRange { int first, last, depth; };

// std::__unguarded_partition_pivot on dst[first, last), by the whole CTA, with the same result as the serial
// scan (gs_partition).  The serial loop alternates "advance lo to the next element that is not before the
// pivot" and "retreat hi to the next element that is not after it", swapping the two until they cross.  Those
// stops are properties of the ORIGINAL array as long as the scans have not met: the k-th swap exchanges the
// k-th such element from the left with the k-th from the right.  So: list both kinds of stops (two CTA-wide
// compactions into `scratch`, 2 x (last - first) ints), count the K leading pairs with l_k < r_k, swap them
// in parallel, and the cut is min(l_(K+1), r_K) -- once hi has passed, everything from r_K on stops lo.
__device__ int cta_partition(HPoint* __restrict__ dst, int first, int last, int* __restrict__ scratch, int* s_warp, int* s_misc) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = last - first;
  if (tid == 0) gs_median_to_first(HpLess(), dst, first, last);
  __syncthreads();
  const HPoint pivot = dst[first];
  int* Ls = scratch;        // ascending positions j in [first + 1, last) with !(dst[j] before pivot)
  int* Rs = scratch + n;    // ascending positions j in [first, last - 1] with !(pivot before dst[j])
  if (tid == 0) { s_misc[0] = 0; s_misc[1] = 0; s_misc[2] = 0; }
  __syncthreads();
  for (int base = first; base < last; base += kSortThreads) {
    const int j = base + tid;
    bool isL = false, isR = false;
    if (j < last) {
      const HPoint e = dst[j];
      isL = j > first && !hp_before(e, pivot);
      isR = !hp_before(pivot, e);
    }
    const unsigned bl = __ballot_sync(0xffffffffu, isL), br = __ballot_sync(0xffffffffu, isR);
    if (lane == 0) { s_warp[warp] = __popc(bl); s_warp[8 + warp] = __popc(br); }
    __syncthreads();
    int offL = s_misc[0], offR = s_misc[1];
    for (int w = 0; w < warp; ++w) { offL += s_warp[w]; offR += s_warp[8 + w]; }
    if (isL) Ls[offL + __popc(bl & ((1u << lane) - 1u))] = j;
    if (isR) Rs[offR + __popc(br & ((1u << lane) - 1u))] = j;
    __syncthreads();
    if (tid == 0) { int a = 0, b = 0; for (int w = 0; w < kSortThreads / 32; ++w) { a += s_warp[w]; b += s_warp[8 + w]; } s_misc[0] += a; s_misc[1] += b; }
    __syncthreads();
  }
  const int nL = s_misc[0], nR = s_misc[1];
  // K = number of leading pairs (k-th stop from the left, k-th from the right) that have not crossed
  int mine = 0;
  for (int k = tid; k < min(nL, nR); k += kSortThreads) mine += Ls[k] < Rs[nR - 1 - k] ? 1 : 0;
  if (mine) atomicAdd(&s_misc[2], mine);
  __syncthreads();
  const int K = s_misc[2];
  for (int k = tid; k < K; k += kSortThreads) {
    const int l = Ls[k], r = Rs[nR - 1 - k];
    const HPoint t = dst[l]; dst[l] = dst[r]; dst[r] = t;
  }
  int cut;
  if (K == 0) cut = Ls[0];
  else { const int rK = Rs[nR - K]; cut = K < nL ? min(Ls[K], rK) : rK; }
  __syncthreads();
  return cut;
}

__global__ void __launch_bounds__(kSortThreads)
harris_sort_kernel(int n_layers, const int* __restrict__ layer_start, const HPoint* __restrict__ pts,
                   const uint8_t* __restrict__ keep, HPoint* __restrict__ sorted, int* __restrict__ layer_kept,
                   HPoint* __restrict__ scratch_all, int cap) {
  __shared__ SortRange s_ranges[2][kSortRanges];
  __shared__ SortRange s_stack[kCoopStack];
  __shared__ int s_count[2];
  __shared__ int s_warp[16];
  __shared__ int s_misc[4];
  __shared__ int s_base, s_sp;
  const int layer = blockIdx.x, frame = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int* ls = layer_start + (long long)frame * (kMaxLayers + 1);
  const int begin = min(ls[layer], cap), end = min(ls[layer + 1], cap);
  const HPoint* src = pts + (long long)frame * cap;
  const uint8_t* kf = keep + (long long)frame * cap;
  HPoint* dst = sorted + (long long)frame * cap + begin;
  // 8 bytes per element of the layer's segment of the survivor buffer, which is written only later
  int* scratch = reinterpret_cast<int*>(scratch_all + (long long)frame * cap + begin);
  if (tid == 0) s_base = 0;
  __syncthreads();
  // stable compaction, kSortThreads candidates per round
  for (int k0 = begin; k0 < end; k0 += kSortThreads) {
    const int k = k0 + tid;
    const bool on = k < end && kf[k];
    const unsigned bal = __ballot_sync(0xffffffffu, on);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int off = s_base;
    for (int w = 0; w < warp; ++w) off += s_warp[w];
    if (on) dst[off + __popc(bal & ((1u << lane) - 1u))] = src[k];
    __syncthreads();
    if (tid == 0) { int t = 0; for (int w = 0; w < kSortThreads / 32; ++w) t += s_warp[w]; s_base += t; }
    __syncthreads();
  }
  const int m = s_base;
  if (tid == 0) {
    layer_kept[frame * kMaxLayers + layer] = m;
    s_count[0] = 0; s_count[1] = 0; s_sp = 0;
    if (m > 16) { s_stack[0] = SortRange{0, m, gcc_depth_limit(m)}; s_sp = 1; }
    else if (m > 1) hp_insertion_sort(dst, dst + m);
  }
  __syncthreads();
  // phase A: ranges above kCoopMin, one at a time, by the whole CTA
  while (s_sp > 0) {
    const SortRange rg = s_stack[s_sp - 1];
    __syncthreads();
    if (tid == 0) --s_sp;
    if (rg.last - rg.first <= kCoopMin || rg.depth == 0) {
      // small enough for one thread (or out of depth: heapsort): hand it to phase B
      if (tid == 0) { const int slot = s_count[0]++; if (slot < kSortRanges) s_ranges[0][slot] = rg; else gcc_sort_range(dst, rg.first, rg.last, rg.depth); }
      __syncthreads();
      continue;
    }
    __syncthreads();
    const int cut = cta_partition(dst, rg.first, rg.last, scratch, s_warp, s_misc);
    if (tid == 0) {
#pragma unroll
      for (int side = 0; side < 2; ++side) {
        const int f = side ? cut : rg.first, l = side ? rg.last : cut;
        if (l - f <= 16) { hp_insertion_sort(dst + f, dst + l); continue; }
        s_stack[s_sp++] = SortRange{f, l, rg.depth - 1};
      }
    }
    __syncthreads();
  }
  if (tid == 0) s_count[0] = min(s_count[0], kSortRanges);
  __syncthreads();
  // phase B: level by level, one thread per open range
  int cur = 0;
  while (s_count[cur] > 0) {
    const int count = s_count[cur];
    for (int r = tid; r < count; r += kSortThreads) {
      const SortRange rg = s_ranges[cur][r];
      if (rg.depth == 0) { hp_heapsort(dst + rg.first, rg.last - rg.first); continue; }
      const int cut = gcc_partition(dst, rg.first, rg.last);
#pragma unroll
      for (int side = 0; side < 2; ++side) {
        const int f = side ? cut : rg.first, l = side ? rg.last : cut;
        if (l - f <= 16) { hp_insertion_sort(dst + f, dst + l); continue; }
        const int slot = atomicAdd(&s_count[cur ^ 1], 1);
        if (slot < kSortRanges) s_ranges[cur ^ 1][slot] = SortRange{f, l, rg.depth - 1};
        else gcc_sort_range(dst, f, l, rg.depth - 1);
      }
    }
    __syncthreads();
    if (tid == 0) { s_count[cur] = 0; s_count[cur ^ 1] = min(s_count[cur ^ 1], kSortRanges); }
    __syncthreads();
    cur ^= 1;
  }
}

struct UniformityGeom {
  long long occ_frame_bytes;
  long long occ_off[kMaxLayers];
  int occ_w[kMaxLayers], occ_h[kMaxLayers];
  float scaling;
};

// One CTA per (frame, layer): EnforceKeyPointUniformity (uniformity-enforcement-inl.h:44-194).
__global__ void __launch_bounds__(256)
harris_uniformity_kernel(PyramidGeom g, UniformityGeom ug, const int* __restrict__ layer_start, const HPoint* __restrict__ sorted,
                         const int* __restrict__ layer_kept, uint8_t* __restrict__ occ_all, HPoint* __restrict__ surv,
                         int* __restrict__ layer_surv, int cap, long long max_kpt) {
  const int frame = blockIdx.y, layer = blockIdx.x, tid = threadIdx.x;
  const int* ls = layer_start + (long long)frame * (kMaxLayers + 1);
  const int begin = min(ls[layer], cap);
  const int m = layer_kept[frame * kMaxLayers + layer];
  const HPoint* pts = sorted + (long long)frame * cap + begin;
  HPoint* out = surv + (long long)frame * cap + begin;
  uint8_t* occ = occ_all + (long long)frame * ug.occ_frame_bytes + ug.occ_off[layer];
  const int ow = ug.occ_w[layer], oh = ug.occ_h[layer];
  {  // every layer's region starts 256-byte aligned and is padded to a multiple of 256 bytes
    uint4* o4 = reinterpret_cast<uint4*>(occ);
    for (long long i = tid; i < ((long long)ow * oh + 15) / 16; i += blockDim.x) o4[i] = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
  __shared__ int s_first[8];
  __shared__ HPoint s_pt;
  __shared__ float s_nsc1;
  int kept = 0;
  bool full = false;
  const float max_score = m > 0 ? (float)pts[0].score : 1.0f;
  // The skip test of a point must see the stamps of all its predecessors.  Occupancy only grows, so a
  // "skip" verdict is final; the points are judged 256 at a time, the first survivor of the batch
  // stamps, and only the points after it are judged again.
  for (int base = 0; base < m && !full; base += 256) {
    const int k = base + tid;
    HPoint p = {0, 0, 0};
    int cell = 0;
    float nsc1 = 0.0f;
    bool undecided = k < m;
    if (undecided) {
      p = pts[k];
      const int cy = (int)((float)(int)p.y * ug.scaling + 16.0f), cx = (int)((float)(int)p.x * ug.scaling + 16.0f);
      cell = cy * ow + cx;
      nsc1 = uniformity_nsc1(p.score, max_score);
    }
    for (;;) {
      if (undecided && (double)nsc1 < (double)*(volatile uint8_t*)(occ + cell)) undecided = false;
      const unsigned bal = __ballot_sync(0xffffffffu, undecided);
      if ((tid & 31) == 0) s_first[tid >> 5] = bal ? (tid | (__ffs(bal) - 1)) : 0x7fffffff;
      __syncthreads();
      int first = s_first[0];
#pragma unroll
      for (int w = 1; w < 8; ++w) first = min(first, s_first[w]);
      if (first == 0x7fffffff) { __syncthreads(); break; }
      if (tid == first) { s_pt = p; s_nsc1 = nsc1; undecided = false; }
      __syncthreads();
      const HPoint sp = s_pt;
      const float nsc = 0.99f * s_nsc1;
      const int cy = (int)((float)(int)sp.y * ug.scaling + 16.0f), cx = (int)((float)(int)sp.x * ug.scaling + 16.0f);
      for (int c = tid; c < 31 * 31; c += blockDim.x) {
        const int y = c / 31, x = c - y * 31;
        uint8_t* o = occ + (long long)(cy + y - 15) * ow + cx + x - 15;
        const int v = (int)*o + uniformity_stamp(x, y, nsc);
        *o = (uint8_t)(v > 255 ? 255 : v);
      }
      if (tid == 0) out[kept] = sp;
      ++kept;
      __syncthreads();
      if ((long long)kept == max_kpt) { full = true; break; }
    }
  }
  if (tid == 0) layer_surv[frame * kMaxLayers + layer] = kept;
}

// One warp per (frame, layer): KeyPointBucketing (key-point-bucketing-inl.h:44-112, key-point-bucketing.h:45-93) with the
// 4 x 4 buckets of ScaleSpaceLayer (scale-space-layer.h:73-74).  The sorted points are taken 32 at a time; a point is kept
// while fewer than maxNumKpt / 16 earlier points of its bucket were kept, i.e. while (bucket count so far + its rank among
// the lanes of the same bucket) is below the quota.  Output keeps the sorted order.
__global__ void __launch_bounds__(32)
harris_bucketing_kernel(PyramidGeom g, const int* __restrict__ layer_start, const HPoint* __restrict__ sorted,
                        const int* __restrict__ layer_kept, HPoint* __restrict__ surv, int* __restrict__ layer_surv, int cap,
                        long long max_kpt) {
  __shared__ unsigned s_cnt[16];
  const int frame = blockIdx.y, layer = blockIdx.x, lane = threadIdx.x;
  const int* ls = layer_start + (long long)frame * (kMaxLayers + 1);
  const int begin = min(ls[layer], cap);
  const int m = layer_kept[frame * kMaxLayers + layer];
  const HPoint* pts = sorted + (long long)frame * cap + begin;
  HPoint* out = surv + (long long)frame * cap + begin;
  const unsigned step_u = 1u + (unsigned)(g.L[layer].w - 1) / 4u, step_v = 1u + (unsigned)(g.L[layer].h - 1) / 4u;
  const unsigned quota = (unsigned)((unsigned long long)max_kpt / 16ull);  // unsigned int _maxNumKeyPointsPerBucket
  if (lane < 16) s_cnt[lane] = 0;
  __syncwarp();
  int kept = 0;
  for (int base = 0; base < m; base += 32) {
    const int k = base + lane;
    const bool valid = k < m;
    HPoint p = {0, 0, 0};
    if (valid) p = pts[k];
    const unsigned b = valid ? (p.x / step_u) * 4u + p.y / step_v : 16u + lane;
    const unsigned same = __match_any_sync(0xffffffffu, b);
    const unsigned rank = __popc(same & ((1u << lane) - 1u));
    const unsigned before = valid ? s_cnt[b] : 0u;
    const bool keep = valid && before + rank < quota;
    __syncwarp();
    if (valid && rank == 0) s_cnt[b] = before + __popc(same);  // counts past the quota change nothing
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (keep) out[kept + __popc(bal & ((1u << lane) - 1u))] = p;
    kept += __popc(bal);
    __syncwarp();
  }
  if (lane == 0) layer_surv[frame * kMaxLayers + layer] = kept;
}

// One CTA per frame: sub-pixel refinement and key-point emission in layer-major order
// (scale-space-layer-inl.h:386-412).
__global__ void __launch_bounds__(256)
harris_emit_kernel(PyramidGeom g, HarrisLayerParams hp, const int* __restrict__ scores, const int* __restrict__ layer_start,
                   const HPoint* __restrict__ surv, const int* __restrict__ layer_surv, int cap, KeyPoint* __restrict__ out,
                   int* __restrict__ counts, int kp_cap) {
  const int frame = blockIdx.x, tid = threadIdx.x;
  const int* ls = layer_start + (long long)frame * (kMaxLayers + 1);
  int base = 0;
  for (int layer = 0; layer < g.n_layers; ++layer) {
    const int m = layer_surv[frame * kMaxLayers + layer];
    const HPoint* pts = surv + (long long)frame * cap + min(ls[layer], cap);
    const LayerGeom& L = g.L[layer];
    const int* sc = scores + (long long)frame * g.frame_elems + L.off;
    for (int k = tid; k < m; k += blockDim.x) {
      const HPoint p = pts[k];
      const int* c = sc + (long long)p.y * L.pitch + p.x;
      float dx, dy;
      harris_subpixel2d((double)c[-L.pitch - 1], (double)c[-L.pitch], (double)c[-L.pitch + 1], (double)c[-1], (double)c[0], (double)c[1],
                        (double)c[L.pitch - 1], (double)c[L.pitch], (double)c[L.pitch + 1], &dx, &dy);
      KeyPoint kp;
      kp.x = (float)(hp.scale[layer] * ((double)((float)(int)p.x + dx) + hp.offset[layer]));
      kp.y = (float)(hp.scale[layer] * ((double)((float)(int)p.y + dy) + hp.offset[layer]));
      kp.size = (float)(hp.scale[layer] * 12.0);
      kp.angle = -1.0f; kp.response = (float)p.score; kp.octave = layer / 2; kp.class_id = -1;
      if (base + k < kp_cap) out[(long long)frame * kp_cap + base + k] = kp;
    }
    base += m;
  }
  if (tid == 0) counts[frame] = base;
}

// HarrisScoreCalculator::InitializeScores + Get2dMaxima on every layer (harris-score-calculator.cc:53-106): score planes and
// the raster-ordered lists of the 2-D maxima (hw.pts, layer-major; hw.det.layer_start[frame][l] = first slot of layer l).
cudaError_t launch_harris_score_maxima(const PyramidGeom& g, const HarrisWorkspace& hw, int n_frames, int thr, int* overflow_flag,
                                       cudaStream_t stream) {
  for (int l = 0; l < g.n_layers; ++l) {
    const LayerGeom& L = g.L[l];
    dim3 grid((L.w + kHsTW - 1) / kHsTW, (L.h + kHsTH - 1) / kHsTH, n_frames);
    harris_scores_kernel<<<grid, kHsThreads, 0, stream>>>(L, g.frame_elems, hw.det.pyr, hw.scores);
  }
  for (int l = 0; l < g.n_layers; ++l) {
    const LayerGeom& L = g.L[l];
    dim3 grid((L.h + 7) / 8, n_frames);
    harris_maxima_kernel<false><<<grid, 256, 0, stream>>>(L, g.frame_elems, hw.scores, thr, hw.det.rowcnt, hw.det.total_rows, hw.det.row_off[l], hw.pts, hw.det.corner_cap);
  }
  cudaError_t e = launch_row_scan(g, hw.det, n_frames, overflow_flag, stream);
  if (e != cudaSuccess) return e;
  for (int l = 0; l < g.n_layers; ++l) {
    const LayerGeom& L = g.L[l];
    dim3 grid((L.h + 7) / 8, n_frames);
    harris_maxima_kernel<true><<<grid, 256, 0, stream>>>(L, g.frame_elems, hw.scores, thr, hw.det.rowcnt, hw.det.total_rows, hw.det.row_off[l], hw.pts, hw.det.corner_cap);
  }
  return cudaGetLastError();
}

cudaError_t launch_harris_detect(const PyramidGeom& g, const HarrisWorkspace& hw, int n_frames, double radius, double abs_thr,
                                 long long max_kpt, KeyPoint* out, int* counts, int kp_cap, int* overflow_flag, cudaStream_t stream) {
  // layer transforms of ScaleSpaceLayer::Create (scale-space-layer-inl.h:60-182)
  HarrisLayerParams hp;
  for (int i = 0; i < g.n_layers; ++i) {
    const bool octave = (i % 2) == 0;
    if (octave) { hp.offset_above[i] = -0.25; hp.offset_below[i] = 1.0 / 6.0; hp.scale_above[i] = 2.0 / 3.0; hp.scale_below[i] = 4.0 / 3.0; hp.scale[i] = i == 0 ? 1.0 : pow(2.0, (double)(i / 2)); }
    else { hp.offset_above[i] = -1.0 / 6.0; hp.offset_below[i] = 0.125; hp.scale_above[i] = 0.75; hp.scale_below[i] = 1.5; hp.scale[i] = pow(2.0, (double)(i / 2)) * 1.5; }
    hp.offset[i] = i == 0 ? 0.0 : hp.scale[i] * 0.5 - 0.5;
  }
  const int thr = (int)abs_thr;
  cudaError_t e = launch_harris_score_maxima(g, hw, n_frames, thr, overflow_flag, stream);
  if (e != cudaSuccess) return e;
  dim3 gp((hw.det.corner_cap + 127) / 128, n_frames);
  harris_nms3d_kernel<<<gp, 128, 0, stream>>>(g, hp, hw.scores, hw.det.layer_start, hw.pts, hw.keep, hw.det.corner_cap, thr);
  harris_sort_kernel<<<dim3(g.n_layers, n_frames), kSortThreads, 0, stream>>>(g.n_layers, hw.det.layer_start, hw.pts, hw.keep, hw.sorted, hw.layer_kept, hw.surv, hw.det.corner_cap);
  if (!(radius > 0.0)) {
    harris_bucketing_kernel<<<dim3(g.n_layers, n_frames), 32, 0, stream>>>(g, hw.det.layer_start, hw.sorted, hw.layer_kept, hw.surv, hw.layer_surv, hw.det.corner_cap, max_kpt);
    harris_emit_kernel<<<n_frames, 256, 0, stream>>>(g, hp, hw.scores, hw.det.layer_start, hw.surv, hw.layer_surv, hw.det.corner_cap, out, counts, kp_cap);
    return cudaGetLastError();
  }
  UniformityGeom ug;
  ug.occ_frame_bytes = hw.occ_frame_bytes;
  ug.scaling = (float)(15.0 / (double)(float)radius);
  for (int l = 0; l < g.n_layers; ++l) { ug.occ_off[l] = hw.occ_off[l]; ug.occ_w[l] = hw.occ_w[l]; ug.occ_h[l] = hw.occ_h[l]; }
  harris_uniformity_kernel<<<dim3(g.n_layers, n_frames), 256, 0, stream>>>(g, ug, hw.det.layer_start, hw.sorted, hw.layer_kept, hw.occ, hw.surv, hw.layer_surv, hw.det.corner_cap, max_kpt);
  harris_emit_kernel<<<n_frames, 256, 0, stream>>>(g, hp, hw.scores, hw.det.layer_start, hw.surv, hw.layer_surv, hw.det.corner_cap, out, counts, kp_cap);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Legacy single-scale brisk::HarrisFeatureDetector (reference brisk/src/harris-feature-detector.cc:56-409,
// brisk/src/vectorized-filters.cc:54-123).  Same building blocks as the scale-space detector with different constants and
// a few quirks that are part of its results: the covariances lose 4 bits BEFORE the (unnormalised, 16-bit wrapping)
// binomial smoothing; the response map is shifted by one pixel against the smoothed planes (CornerHarris writes (i, j)
// from (i, j)); NonmaxSuppress reports x one column to the right of the maximum and uses the fixed threshold 64; the
// uniformity step sorts by the FLOAT response, indexes its half-resolution occupancy map with x as the row, and tests
// the int16 cast of the response.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kHsThreads)
harris_legacy_scores_kernel(LayerGeom L, long long frame_elems, const uint8_t* __restrict__ pyr, int* __restrict__ scores) {
  __shared__ uint8_t s_img[kHsTH + 4][kHsTW + 4];
  __shared__ short s_xx[kHsTH + 2][kHsTW + 2], s_yy[kHsTH + 2][kHsTW + 2], s_xy[kHsTH + 2][kHsTW + 2];
  const int x0 = blockIdx.x * kHsTW, y0 = blockIdx.y * kHsTH, frame = blockIdx.z, tid = threadIdx.x;
  const uint8_t* img = pyr + (long long)frame * frame_elems + L.off;
  int* out = scores + (long long)frame * frame_elems + L.off;
  for (int i = tid; i < (kHsTH + 4) * (kHsTW + 4); i += kHsThreads) {
    const int r = i / (kHsTW + 4), c = i - r * (kHsTW + 4);
    const int y = y0 - 2 + r, x = x0 - 2 + c;
    s_img[r][c] = (y >= 0 && y < L.h && x >= 0 && x < L.w) ? img[(long long)y * L.pitch + x] : 0;
  }
  __syncthreads();
  // GetCovarEntries (:76-196): covariance planes, zero on the image border
  for (int i = tid; i < (kHsTH + 2) * (kHsTW + 2); i += kHsThreads) {
    const int r = i / (kHsTW + 2), c = i - r * (kHsTW + 2);  // plane (r, c) <-> pixel (y0-1+r, x0-1+c) <-> s_img[r+1][c+1]
    const int y = y0 - 1 + r, x = x0 - 1 + c;
    int xx = 0, yy = 0, xy = 0;
    if (y >= 1 && y < L.h - 1 && x >= 1 && x < L.w - 1) {
      int p[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) p[k] = s_img[r + k / 3][c + k % 3];
      harris_products(p, &xx, &yy, &xy);   // the same Scharr kernel x 8 and pmulhw products ...
      xx >>= 4; yy >>= 4; xy >>= 4;        // ... then psraw 4 (:146-157)
    }
    s_xx[r][c] = (short)xx; s_yy[r][c] = (short)yy; s_xy[r][c] = (short)xy;
  }
  __syncthreads();
  for (int i = tid; i < kHsTH * kHsTW; i += kHsThreads) {
    const int r = i / kHsTW, c = i - r * kHsTW;
    const int y = y0 + r, x = x0 + c;
    if (x >= L.w || y >= L.h) continue;
    int v = 0;
    // CornerHarris (:198-268) covers rows [0, h-2) x columns [0, w-2) and reads the smoothed planes AT (y, x), which
    // FilterGauss3by316S (vectorized-filters.cc:54-123) leaves zero on row 0 / column 0
    if (y < L.h - 2 && x < L.w - 2 && y >= 1 && x >= 1) {
      int g[3];
#pragma unroll
      for (int pl = 0; pl < 3; ++pl) {
        const short (*q)[kHsTW + 2] = pl == 0 ? s_xx : (pl == 1 ? s_yy : s_xy);
        const int sum = 4 * q[r + 1][c + 1] + 2 * (q[r][c + 1] + q[r + 2][c + 1] + q[r + 1][c] + q[r + 1][c + 2]) + q[r][c] + q[r][c + 2] +
                        q[r + 2][c] + q[r + 2][c + 2];
        g[pl] = (short)sum;   // paddw wraps
      }
      const int tq = (short)((short)((g[0] >> 1) + (g[1] >> 1)) >> 1);
      v = g[0] * g[1] - g[2] * g[2] - tq * tq;
    }
    out[(long long)y * L.pitch + x] = v;
  }
}

// NonmaxSuppress (:270-322), warp per row.  Column -1 of a row is the last element of the row before it in the
// reference's contiguous map -- a column CornerHarris never writes, i.e. 0.
template <bool FILL>
__global__ void __launch_bounds__(256)
harris_legacy_maxima_kernel(LayerGeom L, long long frame_elems, const int* __restrict__ scores, int* __restrict__ rowcnt,
                            int total_rows, HPoint* __restrict__ pts, int cap) {
  const int frame = blockIdx.y, lane = threadIdx.x & 31;
  const int y = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (y >= L.h) return;
  int* rc = rowcnt + (long long)frame * total_rows + y;
  if (y < 2 || y >= L.h - 2) { if (!FILL && lane == 0) *rc = 0; return; }
  const int* row = scores + (long long)frame * frame_elems + L.off + (long long)y * L.pitch;
  int slot = FILL ? *rc : 0;
  HPoint* out = pts + (long long)frame * cap;
  for (int xb = 0; xb < L.w; xb += 32) {
    const int x = xb + lane;
    bool m = false;
    int c = 0;
    if (x < L.w - 2) {
      c = row[x];
      if (c >= 64) {
        const int* up = row - L.pitch;
        const int* dn = row + L.pitch;
        const int l0 = x ? row[x - 1] : 0, l1 = x ? dn[x - 1] : 0, l2 = x ? up[x - 1] : 0;
        m = !(row[x + 1] > c || l0 > c || dn[x] > c || up[x] > c || dn[x + 1] > c || l1 > c || up[x + 1] > c || l2 > c);
      }
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, m);
    if (FILL && m) {
      const int s = slot + __popc(bal & ((1u << lane) - 1));
      // the key point's response is float(score): sort on the float's value (exact as an int for these magnitudes)
      if (s < cap) out[s] = HPoint{(int)(float)c, (unsigned short)(x + 1), (unsigned short)y};
    }
    slot += __popc(bal);
  }
  if (!FILL && lane == 0) *rc = slot;
}

// EnforceUniformity (:324-393), one CTA per frame, points in std::sort order.
__global__ void __launch_bounds__(256)
harris_legacy_uniformity_kernel(int w, int h, double radius, const HPoint* __restrict__ sorted, const int* __restrict__ layer_kept,
                                uint8_t* __restrict__ occ_all, long long occ_frame_bytes, int cap, KeyPoint* __restrict__ out,
                                int* __restrict__ counts, int kp_cap, int* __restrict__ error_flag) {
  __shared__ float s_lut[31 * 31];
  const int frame = blockIdx.x, tid = threadIdx.x;
  const int m = layer_kept[frame * kMaxLayers];
  const HPoint* pts = sorted + (long long)frame * cap;
  uint8_t* occ = occ_all + (long long)frame * occ_frame_bytes;
  const int H = h / 2 + 32, W = w / 2 + 32;
  for (long long i = tid; i < ((long long)H * W + 15) / 16; i += blockDim.x) reinterpret_cast<uint4*>(occ)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < 31 * 31; i += blockDim.x) {
    // SetRadius (:57-72): the mask is centred at (radius / 2, radius / 2) of the 31 x 31 table
    const int x = i % 31, y = i / 31;
    const double v = 1 - (double)((radius / 2.0 - x) * (radius / 2.0 - x) + (radius / 2.0 - y) * (radius / 2.0 - y)) /
                             (double)(radius / 2.0 * radius / 2.0);
    s_lut[i] = (float)(v > 0.0 ? v : 0.0);
  }
  __syncthreads();
  KeyPoint* dst = out + (long long)frame * kp_cap;
  int kept = 0;
  for (int k = 0; k < m; ++k) {
    const HPoint p = pts[k];
    const float response = (float)p.score;
    // (sic) x selects the ROW of the occupancy map, y the column
    const int cy = (int)((float)(int)p.x / 2 + 16), cx = (int)((float)(int)p.y / 2 + 16);
    if ((long long)(cy + 15) * W + cx + 16 >= (long long)H * W) {   // the reference's 16-byte accesses would leave the map
      if (tid == 0) atomicExch(error_flag, 6);
      break;
    }
    const double s0 = (double)occ[(long long)cy * W + cx];
    const double s1 = s0 * s0;
    const short r16 = (short)(int)response;   // static_cast<int16_t>(float): cvttss2si, low 16 bits
    __syncthreads();                          // every thread has read the test byte before anybody stamps
    if ((double)r16 < s1 * s1) continue;
    const float nsc = (float)sqrt(sqrt((double)response));
    for (int c = tid; c < 31 * 31; c += blockDim.x) {
      const int y = c / 31, x = c - y * 31;
      uint8_t* o = occ + (long long)(cy + y - 15) * W + cx + x - 15;
      const int v = (int)*o + ((int)(s_lut[c] * nsc) & 0xff);   // (char)(float * float), then paddusb
      *o = (uint8_t)(v > 255 ? 255 : v);
    }
    if (tid == 0 && kept < kp_cap) dst[kept] = KeyPoint{(float)(int)p.x, (float)(int)p.y, 10.0f, -1.0f, response, 0, -1};
    ++kept;
    __syncthreads();
  }
  if (tid == 0) counts[frame] = kept;
}

// hw: scores, pts, keep, sorted, layer_kept, surv (sort scratch), occ with occ_frame_bytes >= (h/2+32) * (w/2+32) + 16.
cudaError_t launch_harris_legacy_detect(const PyramidGeom& g, const HarrisWorkspace& hw, int n_frames, double radius, KeyPoint* out,
                                        int* counts, int kp_cap, int* error_flag, cudaStream_t stream) {
  const LayerGeom& L = g.L[0];
  dim3 gs((L.w + kHsTW - 1) / kHsTW, (L.h + kHsTH - 1) / kHsTH, n_frames);
  harris_legacy_scores_kernel<<<gs, kHsThreads, 0, stream>>>(L, g.frame_elems, hw.det.pyr, hw.scores);
  dim3 gm((L.h + 7) / 8, n_frames);
  harris_legacy_maxima_kernel<false><<<gm, 256, 0, stream>>>(L, g.frame_elems, hw.scores, hw.det.rowcnt, hw.det.total_rows, hw.pts, hw.det.corner_cap);
  cudaError_t e = launch_row_scan(g, hw.det, n_frames, error_flag, stream);
  if (e != cudaSuccess) return e;
  harris_legacy_maxima_kernel<true><<<gm, 256, 0, stream>>>(L, g.frame_elems, hw.scores, hw.det.rowcnt, hw.det.total_rows, hw.pts, hw.det.corner_cap);
  e = cudaMemsetAsync(hw.keep, 1, (size_t)n_frames * hw.det.corner_cap, stream);
  if (e != cudaSuccess) return e;
  harris_sort_kernel<<<dim3(1, n_frames), kSortThreads, 0, stream>>>(1, hw.det.layer_start, hw.pts, hw.keep, hw.sorted, hw.layer_kept, hw.surv, hw.det.corner_cap);
  harris_legacy_uniformity_kernel<<<n_frames, 256, 0, stream>>>(L.w, L.h, radius, hw.sorted, hw.layer_kept, hw.occ, hw.occ_frame_bytes, hw.det.corner_cap,
                                                                 out, counts, kp_cap, error_flag);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// "Use passed key points" (scale-space-feature-detector.h:103-108, scale-space-layer-inl.h:198-208,370-428):
// detect() on a non-empty vector computes no scores; the points with response > 1e6 are re-filtered by the
// uniformity enforcement (or the bucketing) and come back unrefined.  One layer only (octaves == 0).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
harris_import_kernel(const KeyPoint* __restrict__ in, const int* __restrict__ in_counts, int in_cap, int w, int h,
                     HPoint* __restrict__ pts, uint8_t* __restrict__ keep, int* __restrict__ layer_start, int cap,
                     int* __restrict__ error_flag) {
  const int frame = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_in = min(min(in_counts[frame], in_cap), cap);
  if (j == 0) {
    int* ls = layer_start + (long long)frame * (kMaxLayers + 1);
    ls[0] = 0; ls[1] = n_in;
  }
  if (j >= n_in) return;
  const KeyPoint kp = in[(long long)frame * in_cap + j];
  bool on = (double)kp.response > 1e6;
  HPoint p = {0, 0, 0};
  if (on) {
    // PointWithScore(int score, uint16_t x, uint16_t y) from float fields: defined for values in range only
    if (!(kp.response < 2147483648.0f) || !(kp.x >= 0.0f && kp.x < (float)w && kp.y >= 0.0f && kp.y < (float)h)) {
      atomicExch(error_flag, 4);
      on = false;
    } else {
      p.score = (int)kp.response; p.x = (unsigned short)(int)kp.x; p.y = (unsigned short)(int)kp.y;
    }
  }
  pts[(long long)frame * cap + j] = p;
  keep[(long long)frame * cap + j] = on ? 1 : 0;
}

__global__ void __launch_bounds__(256)
harris_emit_passed_kernel(const KeyPoint* __restrict__ in, const int* __restrict__ in_counts, int in_cap, const HPoint* __restrict__ surv,
                          const int* __restrict__ layer_kept, const int* __restrict__ layer_surv, int cap, KeyPoint* __restrict__ out,
                          int* __restrict__ counts, int kp_cap) {
  const int frame = blockIdx.x, tid = threadIdx.x;
  if (layer_kept[frame * kMaxLayers] == 0) {
    // no point passed the response test: the reference returns before it clears the vector (:370-371)
    const int n_in = min(in_counts[frame], in_cap);
    for (int k = tid; k < n_in && k < kp_cap; k += blockDim.x) out[(long long)frame * kp_cap + k] = in[(long long)frame * in_cap + k];
    if (tid == 0) counts[frame] = n_in;
    return;
  }
  const int m = layer_surv[frame * kMaxLayers];
  for (int k = tid; k < m && k < kp_cap; k += blockDim.x) {
    const HPoint p = surv[(long long)frame * cap + k];
    KeyPoint kp;
    kp.x = (float)(1.0 * ((double)(int)p.x + 0.0)); kp.y = (float)(1.0 * ((double)(int)p.y + 0.0));
    kp.size = (float)(1.0 * 12.0); kp.angle = -1.0f; kp.response = (float)p.score; kp.octave = 0; kp.class_id = -1;
    out[(long long)frame * kp_cap + k] = kp;
  }
  if (tid == 0) counts[frame] = m;
}

cudaError_t launch_harris_passed(const PyramidGeom& g, const HarrisWorkspace& hw, int n_frames, double radius, long long max_kpt,
                                 const KeyPoint* in, const int* in_counts, int in_cap, int in_max, KeyPoint* out, int* counts,
                                 int kp_cap, int* error_flag, cudaStream_t stream) {
  if (g.n_layers != 1 || in_max <= 0 || in_max > hw.det.corner_cap) return cudaErrorInvalidValue;
  const int cap = hw.det.corner_cap;
  harris_import_kernel<<<dim3((in_max + 255) / 256, n_frames), 256, 0, stream>>>(in, in_counts, in_cap, g.L[0].w, g.L[0].h, hw.pts, hw.keep,
                                                                                 hw.det.layer_start, cap, error_flag);
  harris_sort_kernel<<<dim3(1, n_frames), kSortThreads, 0, stream>>>(1, hw.det.layer_start, hw.pts, hw.keep, hw.sorted, hw.layer_kept, hw.surv, cap);
  if (!(radius > 0.0)) {
    harris_bucketing_kernel<<<dim3(1, n_frames), 32, 0, stream>>>(g, hw.det.layer_start, hw.sorted, hw.layer_kept, hw.surv, hw.layer_surv, cap, max_kpt);
  } else {
    UniformityGeom ug;
    ug.occ_frame_bytes = hw.occ_frame_bytes;
    ug.scaling = (float)(15.0 / (double)(float)radius);
    for (int l = 0; l < g.n_layers; ++l) { ug.occ_off[l] = hw.occ_off[l]; ug.occ_w[l] = hw.occ_w[l]; ug.occ_h[l] = hw.occ_h[l]; }
    harris_uniformity_kernel<<<dim3(1, n_frames), 256, 0, stream>>>(g, ug, hw.det.layer_start, hw.sorted, hw.layer_kept, hw.occ, hw.surv, hw.layer_surv, cap, max_kpt);
  }
  harris_emit_passed_kernel<<<n_frames, 256, 0, stream>>>(in, in_counts, in_cap, hw.surv, hw.layer_kept, hw.layer_surv, cap, out, counts, kp_cap);
  return cudaGetLastError();
}

}  // namespace briskb200
