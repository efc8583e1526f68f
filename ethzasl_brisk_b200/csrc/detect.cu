// AGAST 9-16 corner detection with the reference's local-contrast threshold map,
// fused into one pass per pyramid layer (sm_100a).
//
// Replaces BriskLayer::CalculateThresholdMap (reference brisk/src/brisk-layer.cc:
// 278-598, three full-image SSE passes) + OastDetector9_16::detect (agast/src/
// oast9-16.cc:43-1859, a generated decision tree) + the score write-back of
// BriskLayer::GetAgastPoints (brisk-layer.cc:99-117).  A CTA stages a tile with
// a 3-pixel halo in shared memory, builds the 3x3 min / max planes there,
// derives the per-pixel threshold T, and runs the 9-of-16 segment test as a
// bit-mask run-length check.  Output: the u16 corner map (T at corners, 0
// elsewhere) and per-row corner counts; `corner_lists` then turns those into
// raster-ordered corner lists (the order the reference's detector emits).
#include <cuda_runtime.h>

#include "brisk_math.cuh"
#include "kernels.h"

namespace briskb200 {

constexpr int kDetTW = 128, kDetTH = 64, kDetThreads = 256;
constexpr int kDetStrip = 32;                 // rows per thread: 2 strips of 32 rows, 128 columns each
constexpr int kDetSW = kDetTW + 8;            // staged row: [x0-4, x0+TW+4), 4-byte aligned
constexpr int kDetSH = kDetTH + 6;            // staged rows: [y0-3, y0+TH+3)

// The reference's threshold map (brisk-layer.cc:278-598) is max - min over the centre, four diagonals
// and four 3x3 blocks; that footprint is exactly the 37-pixel disk with row half-widths
// 1,2,3,3,3,2,1 (it contains the 16-pixel FAST ring, which is why a corner's score equals its
// threshold-map value).  Phase 1: each thread walks down one column; per row it forms the horizontal
// max / min over widths 3, 5 and 7 from seven shared-memory bytes and folds them into seven rolling
// per-output-row accumulators held in registers (no min/max planes are materialised), then runs the
// cheap 4-point compass pre-test and queues the survivors.  Phase 2: the queue is processed densely,
// one candidate per thread, with the full 9-of-16 run test -- so the expensive test only costs the
// lanes that need it.
__global__ void __launch_bounds__(kDetThreads, 3)
agast_detect_kernel(LayerGeom L, long long frame_elems, const uint8_t* __restrict__ pyr, uint16_t* __restrict__ cm,
                    int* __restrict__ rowcnt, int total_rows, int row_off, int thresh) {
  __shared__ __align__(16) uint8_t s_img[kDetSH][kDetSW];
  __shared__ uint32_t s_queue[kDetTW * kDetTH];  // cx | ry << 8 | T << 16
  __shared__ int s_count;
  __shared__ int s_rows[kDetTH];

  const int x0 = blockIdx.x * kDetTW, y0 = blockIdx.y * kDetTH, frame = blockIdx.z;
  const uint8_t* img = pyr + (long long)frame * frame_elems + L.off;
  uint16_t* cmap = cm + (long long)frame * frame_elems + L.off;
  const int tid = threadIdx.x;

  // stage tile + halo (zero outside the image; such pixels never reach a valid output)
  for (int i = tid; i < kDetSH * (kDetSW / 4); i += kDetThreads) {
    const int r = i / (kDetSW / 4), c = i - r * (kDetSW / 4);
    const int y = y0 - 3 + r, x = x0 - 4 + 4 * c;
    uint32_t v = 0;
    if (y >= 0 && y < L.h && x >= 0 && x < L.pitch) v = *reinterpret_cast<const uint32_t*>(img + (long long)y * L.pitch + x);
    *reinterpret_cast<uint32_t*>(&s_img[r][4 * c]) = v;
  }
  if (tid == 0) s_count = 0;
  if (tid < kDetTH) s_rows[tid] = 0;
  __syncthreads();

  const int cmp = (thresh * kLowerThreshold) / 100;  // ast-detector.h:62-68
  {
    const int cx = tid & (kDetTW - 1), strip = tid >> 7;
    const int x = x0 + cx;
    const int ys = y0 + strip * kDetStrip;  // first output row of this thread
    const int sc = cx + 4;                  // staged column of x
    int hi[7], lo[7];                       // rolling accumulators, slot = (output row - (ys - 6)) % 7
#pragma unroll
    for (int k = 0; k < 7; ++k) { hi[k] = 0; lo[k] = 255; }
    if (ys < L.h) {
#pragma unroll 7
      for (int i = 0; i < 42; ++i) {  // 42 = 6 groups of 7 >= kDetStrip + 6; slots are static after unrolling by 7
        if (i >= kDetStrip + 6) break;
        // source row r = ys - 3 + i  <->  staged row (ys - y0) + i
        const uint8_t* row = &s_img[strip * kDetStrip + i][sc - 3];
        const int a0 = row[0], a1 = row[1], a2 = row[2], a3 = row[3], a4 = row[4], a5 = row[5], a6 = row[6];
        const int M3 = imax(imax(a2, a3), a4), M5 = imax(imax(M3, a1), a5), M7 = imax(imax(M5, a0), a6);
        const int m3 = imin(imin(a2, a3), a4), m5 = imin(imin(m3, a1), a5), m7 = imin(imin(m5, a0), a6);
        hi[(i + 6) % 7] = M3;                       lo[(i + 6) % 7] = m3;   // output row r+3 starts here
        hi[(i + 5) % 7] = imax(hi[(i + 5) % 7], M5); lo[(i + 5) % 7] = imin(lo[(i + 5) % 7], m5);
        hi[(i + 4) % 7] = imax(hi[(i + 4) % 7], M7); lo[(i + 4) % 7] = imin(lo[(i + 4) % 7], m7);
        hi[(i + 3) % 7] = imax(hi[(i + 3) % 7], M7); lo[(i + 3) % 7] = imin(lo[(i + 3) % 7], m7);
        hi[(i + 2) % 7] = imax(hi[(i + 2) % 7], M7); lo[(i + 2) % 7] = imin(lo[(i + 2) % 7], m7);
        hi[(i + 1) % 7] = imax(hi[(i + 1) % 7], M5); lo[(i + 1) % 7] = imin(lo[(i + 1) % 7], m5);
        hi[i % 7] = imax(hi[i % 7], M3);             lo[i % 7] = imin(lo[i % 7], m3);   // output row r-3 complete
        if (i < 6) continue;
        const int ry = strip * kDetStrip + i - 6;  // row inside the tile
        const int y = y0 + ry;
        if (y >= L.h) break;
        const int T = hi[i % 7] - lo[i % 7];
        if (x < L.w) cmap[(long long)y * L.pitch + x] = 0;  // corners are written in phase 2
        if (T >= cmp && x >= 3 && x < L.w - 3 && y >= 3 && y < L.h - 3) {
          const int sr = ry + 3;  // staged row of y
          const int t = T < kLowerThreshold ? kLowerThreshold : (T > kUpperThreshold ? kUpperThreshold : T);
          const int b2 = (t * thresh) / 100;
          const int c = s_img[sr][sc], cb = c + b2, c_b = c - b2;
          // compass points first: a 9-arc holds at least two of them
          const int p0 = s_img[sr][sc - 3], p4 = s_img[sr - 3][sc], p8 = s_img[sr][sc + 3], p12 = s_img[sr + 3][sc];
          const int nb = (p0 > cb) + (p4 > cb) + (p8 > cb) + (p12 > cb);
          const int nd = (p0 < c_b) + (p4 < c_b) + (p8 < c_b) + (p12 < c_b);
          if (nb >= 2 || nd >= 2) s_queue[atomicAdd(&s_count, 1)] = (uint32_t)cx | ((uint32_t)ry << 8) | ((uint32_t)T << 16);
        }
      }
    }
  }
  __syncthreads();
  // phase 2: full segment test on the queued candidates
  const int n_cand = s_count;
  for (int q = tid; q < n_cand; q += kDetThreads) {
    const uint32_t e = s_queue[q];
    const int cx = e & 0xff, ry = (e >> 8) & 0xff, T = e >> 16;
    const int sr = ry + 3, sc = cx + 4;
    const int t = T < kLowerThreshold ? kLowerThreshold : (T > kUpperThreshold ? kUpperThreshold : T);
    const int b2 = (t * thresh) / 100;
    const int c = s_img[sr][sc], cb = c + b2, c_b = c - b2;
    // ring of agast/include/agast/oast9-16.h:99-116
    int r[16];
    r[0] = s_img[sr][sc - 3];      r[1] = s_img[sr - 1][sc - 3]; r[2] = s_img[sr - 2][sc - 2]; r[3] = s_img[sr - 3][sc - 1];
    r[4] = s_img[sr - 3][sc];      r[5] = s_img[sr - 3][sc + 1]; r[6] = s_img[sr - 2][sc + 2]; r[7] = s_img[sr - 1][sc + 3];
    r[8] = s_img[sr][sc + 3];      r[9] = s_img[sr + 1][sc + 3]; r[10] = s_img[sr + 2][sc + 2]; r[11] = s_img[sr + 3][sc + 1];
    r[12] = s_img[sr + 3][sc];     r[13] = s_img[sr + 3][sc - 1]; r[14] = s_img[sr + 2][sc - 2]; r[15] = s_img[sr + 1][sc - 3];
    uint32_t mb = 0, md = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) { mb |= (uint32_t)(r[k] > cb) << k; md |= (uint32_t)(r[k] < c_b) << k; }
    // 9 contiguous set bits on the circular 16-bit mask
    mb |= mb << 16; md |= md << 16;
    uint32_t xb = mb & (mb >> 1); xb &= xb >> 2; xb &= xb >> 4; xb &= mb >> 8;
    uint32_t xd = md & (md >> 1); xd &= xd >> 2; xd &= xd >> 4; xd &= md >> 8;
    if (((xb | xd) & 0xffffu) != 0) {
      cmap[(long long)(y0 + ry) * L.pitch + x0 + cx] = (uint16_t)T;
      atomicAdd(&s_rows[ry], 1);
    }
  }
  __syncthreads();
  if (tid < kDetTH && s_rows[tid] && y0 + tid < L.h) atomicAdd(&rowcnt[(long long)frame * total_rows + row_off + y0 + tid], s_rows[tid]);
}

cudaError_t launch_agast_detect(const PyramidGeom& g, const DetectWorkspace& ws, int n_frames, int thresh, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(ws.rowcnt, 0, sizeof(int) * (size_t)n_frames * ws.total_rows, stream);
  if (e != cudaSuccess) return e;
  for (int l = 0; l < g.n_layers; ++l) {
    const LayerGeom& L = g.L[l];
    dim3 grid((L.w + kDetTW - 1) / kDetTW, (L.h + kDetTH - 1) / kDetTH, n_frames);
    agast_detect_kernel<<<grid, kDetThreads, 0, stream>>>(L, g.frame_elems, ws.pyr, ws.cm, ws.rowcnt, ws.total_rows, ws.row_off[l], thresh);
  }
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Ordered corner lists.
// ---------------------------------------------------------------------------

// One CTA per frame: exclusive prefix sum over the per-row counts of all layers
// (rows concatenated layer by layer), in place; also the first slot of every
// layer and the frame total.
__global__ void __launch_bounds__(1024)
row_scan_kernel(int* __restrict__ rowcnt, int total_rows, PyramidGeom g, DetectWorkspace ws, int* __restrict__ overflow_flag) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  const int frame = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int* rc = rowcnt + (long long)frame * total_rows;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < total_rows; base += 1024) {
    const int i = base + tid;
    const int v = i < total_rows ? rc[i] : 0;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      int w = s_warp[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, d); if (lane >= d) w += t; }
      s_warp[lane] = w;
    }
    __syncthreads();
    const int excl = s_carry + (warp ? s_warp[warp - 1] : 0) + inc - v;
    if (i < total_rows) rc[i] = excl;
    __syncthreads();
    if (tid == 1023) s_carry = excl + v;
    __syncthreads();
  }
  if (tid <= g.n_layers) {
    int* ls = ws.layer_start + (long long)frame * (kMaxLayers + 1);
    ls[tid] = tid < g.n_layers ? rc[ws.row_off[tid]] : s_carry;
    if (tid == g.n_layers && s_carry > ws.corner_cap) atomicExch(overflow_flag, 1);
  }
}

// One warp per row: emit the row's corners at their slots.
__global__ void __launch_bounds__(256)
corner_fill_kernel(LayerGeom L, int layer, long long frame_elems, const uint16_t* __restrict__ cm, const int* __restrict__ rowcnt,
                   int total_rows, int row_off, uint32_t* __restrict__ corners, int corner_cap) {
  const int frame = blockIdx.y, lane = threadIdx.x & 31;
  const int y = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (y < 3 || y >= L.h - 3) return;
  const uint16_t* row = cm + (long long)frame * frame_elems + L.off + (long long)y * L.pitch;
  int slot = rowcnt[(long long)frame * total_rows + row_off + y];
  uint32_t* out = corners + (long long)frame * corner_cap;
  for (int xb = 0; xb < L.w; xb += 64) {
    // two pixels per lane
    const int x = xb + 2 * lane;
    uint32_t v = 0;
    if (x < L.pitch) v = *reinterpret_cast<const uint32_t*>(row + x);
    const bool c0 = (v & 0xffffu) != 0 && x < L.w, c1 = (v >> 16) != 0 && x + 1 < L.w;
    const uint32_t m0 = __ballot_sync(0xffffffffu, c0), m1 = __ballot_sync(0xffffffffu, c1);
    if (m0 | m1) {
      const uint32_t below = (1u << lane) - 1;
      const int rank0 = __popc(m0 & below) + __popc(m1 & below);
      if (c0) { const int s = slot + rank0; if (s < corner_cap) out[s] = (uint32_t)x | ((uint32_t)y << 13) | ((uint32_t)layer << 26); }
      if (c1) { const int s = slot + rank0 + (c0 ? 1 : 0); if (s < corner_cap) out[s] = (uint32_t)(x + 1) | ((uint32_t)y << 13) | ((uint32_t)layer << 26); }
      slot += __popc(m0) + __popc(m1);
    }
  }
}

cudaError_t launch_row_scan(const PyramidGeom& g, const DetectWorkspace& ws, int n_frames, int* overflow_flag, cudaStream_t stream) {
  row_scan_kernel<<<n_frames, 1024, 0, stream>>>(ws.rowcnt, ws.total_rows, g, ws, overflow_flag);
  return cudaGetLastError();
}

cudaError_t launch_corner_lists(const PyramidGeom& g, const DetectWorkspace& ws, int n_frames, int* overflow_flag, cudaStream_t stream) {
  row_scan_kernel<<<n_frames, 1024, 0, stream>>>(ws.rowcnt, ws.total_rows, g, ws, overflow_flag);
  for (int l = 0; l < g.n_layers; ++l) {
    const LayerGeom& L = g.L[l];
    dim3 grid((L.h + 7) / 8, n_frames);
    corner_fill_kernel<<<grid, 256, 0, stream>>>(L, l, g.frame_elems, ws.cm, ws.rowcnt, ws.total_rows, ws.row_off[l], ws.corners, ws.corner_cap);
  }
  return cudaGetLastError();
}

}  // namespace briskb200
