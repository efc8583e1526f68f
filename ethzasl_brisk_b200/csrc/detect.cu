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
#include "fast_packed.cuh"
#include "kernels.h"

namespace briskb200 {

constexpr int kDetTW = 256, kDetTH = 60, kDetThreads = 128;
constexpr int kDetStrip = 30;                 // rows per thread: 2 strips of 30 rows; a thread owns 4 adjacent columns
constexpr int kDetRows = kDetStrip + 6;       // source rows a thread walks (36 = 3 * 12, the unroll period)
constexpr int kDetSW = kDetTW + 8;            // staged row: [x0-4, x0+TW+4), 4-byte aligned
constexpr int kDetSH = kDetTH + 6;            // staged rows: [y0-3, y0+TH+3)
#ifndef BRISK_DET_QUEUE
#define BRISK_DET_QUEUE 6144
#endif
#ifndef BRISK_DET_MINB
#define BRISK_DET_MINB 1
#endif
constexpr int kDetQueue = BRISK_DET_QUEUE;               // candidate queue (cx | ry << 8); beyond it candidates are tested in place
constexpr uint32_t kB2None = 0x3fffu;         // contrast no pixel reaches: "threshold map below the lower bound"

// 9-of-16 segment test of OastDetector9_16::detect on the staged tile (ring of agast/include/agast/oast9-16.h:99-116).
__device__ __forceinline__ bool segment_test(const uint8_t (*s_img)[kDetSW], int sr, int sc, int b2) {
  const int c = s_img[sr][sc], cb = c + b2, c_b = c - b2;
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
  return ((xb | xd) & 0xffffu) != 0;
}

// The same test for TWO pixels at once, on packed pairs (fast_packed.cuh): a 9-arc brighter than c + b exists  <=>  the
// largest arc minimum exceeds c + b; darker alike.  About half the instructions per pixel of the bit-mask form, which
// matters: one pixel in seven to ten reaches this test.  Returns bit 0 / bit 1 = pixel 0 / pixel 1 is a corner.
__device__ __forceinline__ uint32_t segment_test_pair(const uint8_t (*s_img)[kDetSW], int sr0, int sc0, int b0, int sr1, int sc1, int b1) {
  const uint8_t* a = &s_img[sr0][sc0];
  const uint8_t* d = &s_img[sr1][sc1];
  uint32_t p[16];
#define BRISK_RING(i, dx, dy) p[i] = (uint32_t)a[(dy) * kDetSW + (dx)] | ((uint32_t)d[(dy) * kDetSW + (dx)] << 16)
  BRISK_RING(0, -3, 0);  BRISK_RING(1, -3, -1); BRISK_RING(2, -2, -2); BRISK_RING(3, -1, -3);
  BRISK_RING(4, 0, -3);  BRISK_RING(5, 1, -3);  BRISK_RING(6, 2, -2);  BRISK_RING(7, 3, -1);
  BRISK_RING(8, 3, 0);   BRISK_RING(9, 3, 1);   BRISK_RING(10, 2, 2);  BRISK_RING(11, 1, 3);
  BRISK_RING(12, 0, 3);  BRISK_RING(13, -1, 3); BRISK_RING(14, -2, 2); BRISK_RING(15, -3, 1);
#undef BRISK_RING
  uint32_t bright, dark;
  arc_extrema16x2(p, &bright, &dark);
  constexpr uint32_t kHi = 0x80008000u;
  const uint32_t c = (uint32_t)a[0] | ((uint32_t)d[0] << 16), b = (uint32_t)b0 | ((uint32_t)b1 << 16);
  // lane test A > B as ((B | 0x8000) - A) losing bit 15 (all values stay below 0x8000)
  const uint32_t tb = ((c + b) | kHi) - bright;   // bit 15 clear <=> bright > c + b
  const uint32_t td = ((dark + b) | kHi) - c;     // bit 15 clear <=> c > dark + b
  const uint32_t hit = ~(tb & td) & kHi;
  return (hit >> 15 & 1u) | (hit >> 30 & 2u);
}

// The reference's threshold map (brisk-layer.cc:278-598) is max - min over the centre, four diagonals
// and four 3x3 blocks; that footprint is exactly the 37-pixel disk with row half-widths
// 1,2,3,3,3,2,1 (it contains the 16-pixel FAST ring, which is why a corner's score equals its
// threshold-map value).
//
// Phase 1 works on packed pairs of pixels (two 16-bit lanes per register, VIMNMX3.U16x2): a thread owns
// four adjacent columns x..x+3 as the pairs E = (x, x+2) and O = (x+1, x+3) and walks down 30 rows.  Per
// source row it loads 12 staged bytes, spreads them into the pairs P_j = (b_j, b_j+2), forms the
// horizontal max / min over widths 3, 5 and 7 with three 3-input operations each, and combines seven
// rows with three more:  out(y) = max3(G(y-1), M7(y), H(y+3)),  G(r) = max3(M3(r-2), M5(r-1), M7(r)),
// H(r) = max3(M7(r-2), M5(r-1), M3(r)); the short histories live in registers (the loop is unrolled by
// 12, a multiple of every history depth, so all slots are static).  The 4-point compass pre-test (a
// 9-arc holds two adjacent compass points) is done on the pairs as well: two adjacent compass pixels
// brighter than c + b  <=>  min(max(left, right), max(up, down)) exceeds it.  The verdicts are shifted into two 64-bit
// registers (two bits per row) and the threshold-map values go to a byte plane in shared memory; the
// survivors (a few per cent) are queued once the walk is over, so that the row loop has no divergent path.
// Phase 2: the queue is processed densely, one candidate per thread, with the full 9-of-16 run test.
// LOWER: lower bound of the threshold map (BriskScaleSpace::kDefaultLowerThreshold = 10 for detection;
// 0 for the pyramid that BriskFeatureDetector::ComputeScale builds, brisk-feature-detector.cc:90).
template <int LOWER>
__global__ void __launch_bounds__(kDetThreads, BRISK_DET_MINB)
agast_detect_kernel(LayerGeom L, long long frame_elems, const uint8_t* __restrict__ pyr, uint16_t* __restrict__ cm,
                    int* __restrict__ rowcnt, int total_rows, int row_off, int thresh) {
  __shared__ __align__(16) uint8_t s_img[kDetSH][kDetSW];
  __shared__ __align__(4) uint8_t s_T[kDetTH][kDetTW];  // threshold-map values of the tile
  __shared__ uint16_t s_queue[kDetQueue];  // cx | ry << 8
  __shared__ uint16_t s_b2[256];           // threshold-map value -> contrast b of the segment test
  __shared__ int s_count;
  __shared__ int s_rows[kDetTH];

  const int x0 = blockIdx.x * kDetTW, y0 = blockIdx.y * kDetTH, frame = blockIdx.z;
  const uint8_t* img = pyr + (long long)frame * frame_elems + L.off;
  uint16_t* cmap = cm + (long long)frame * frame_elems + L.off;
  const int tid = threadIdx.x;

  // stage tile + halo (zero outside the image; such pixels never reach a valid output) with asynchronous copies: all of a
  // thread's 34 words are in flight at once, and the tables below are set up while they arrive (with register-staged
  // loads, four at a time, a fifth of the kernel's stall samples sat on the stores behind them)
  for (int i = tid; i < kDetSH * (kDetSW / 4); i += kDetThreads) {
    const int r = i / (kDetSW / 4), c = i - r * (kDetSW / 4);
    const int y = y0 - 3 + r, x = x0 - 4 + 4 * c;
    const bool inside = y >= 0 && y < L.h && x >= 0 && x < L.pitch;
    const uint8_t* src = inside ? img + (long long)y * L.pitch + x : img;
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&s_img[r][4 * c]);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(inside ? 4 : 0) : "memory");   // src-size 0: zero fill
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  {
    // ast-detector.h:62-68 and oast9-16.cc:86-100: no corner where T < (thresh * lower) / 100, else
    // b = (clamp(T, lower, upper) * thresh) / 100
    const int cmp = (thresh * LOWER) / 100;
    for (int T = tid; T < 256; T += kDetThreads) {
      const int t = T < LOWER ? LOWER : (T > kUpperThreshold ? kUpperThreshold : T);
      s_b2[T] = (uint16_t)(T >= cmp ? (t * thresh) / 100 : kB2None);
    }
  }
  for (int i = tid; i < kDetQueue / 2; i += kDetThreads) reinterpret_cast<uint32_t*>(s_queue)[i] = 0xffffffffu;
  if (tid == 0) s_count = 0;
  if (tid < kDetTH) s_rows[tid] = 0;
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  {
    const int g = tid & 63, strip = tid >> 6;
    const int cx = 4 * g, x = x0 + cx;
    const int ys = y0 + strip * kDetStrip;  // first output row of this thread
    const int sc = cx + 4;                  // staged column of x
    if (ys < L.h && x < L.pitch) {
      constexpr uint32_t kHi = 0x80008000u;
      // histories, indexed by source row modulo their depth; [0] = max / pair E, see below
      uint32_t m3[4][2], m5[4], m7[4][3], gg[4][4];  // [pair E max, pair E min, pair O max, pair O min]
      uint32_t ce[6], co[6], le[3], lo[3], re[3], ro[3];  // centre / left (x-3) / right (x+3) pixels of the pairs
      unsigned long long acc_lo = 0, acc_hi = 0;          // compass verdicts, two bits per row
      int last_ry = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        m3[k][0] = m3[k][1] = m5[k] = m7[k][0] = m7[k][1] = m7[k][2] = 0;
        gg[k][0] = gg[k][1] = gg[k][2] = gg[k][3] = 0;
      }
#pragma unroll
      for (int k = 0; k < 6; ++k) ce[k] = co[k] = 0;
#pragma unroll
      for (int k = 0; k < 3; ++k) le[k] = lo[k] = re[k] = ro[k] = 0;
#pragma unroll 12
      for (int i = 0; i < kDetRows; ++i) {
        // source row r = ys - 3 + i  <->  staged row (ys - y0) + i; bytes b0..b11 = staged columns sc-4 .. sc+7
        const uint32_t* row = reinterpret_cast<const uint32_t*>(&s_img[strip * kDetStrip + i][sc - 4]);
        const uint32_t w0 = row[0], w1 = row[1], w2 = row[2];
        const uint32_t f0 = __funnelshift_r(w0, w1, 16), f1 = __funnelshift_r(w1, w2, 16);  // b2..b5, b6..b9
        // P_j = (b_j, b_j+2) as two 16-bit lanes
        const uint32_t P1 = __byte_perm(w0, 0, 0x4341), P2 = __byte_perm(f0, 0, 0x4240), P3 = __byte_perm(f0, 0, 0x4341);
        const uint32_t P4 = __byte_perm(w1, 0, 0x4240), P5 = __byte_perm(w1, 0, 0x4341), P6 = __byte_perm(f1, 0, 0x4240);
        const uint32_t P7 = __byte_perm(f1, 0, 0x4341), P8 = __byte_perm(w2, 0, 0x4240);
        // horizontal extrema over widths 3 / 5 / 7: pair E is centred on P4, pair O on P5
        uint32_t M3[4], M5[4], M7[4];
        M3[0] = __vimax3_u16x2(P3, P4, P5); M5[0] = __vimax3_u16x2(M3[0], P2, P6); M7[0] = __vimax3_u16x2(M5[0], P1, P7);
        M3[1] = __vimin3_u16x2(P3, P4, P5); M5[1] = __vimin3_u16x2(M3[1], P2, P6); M7[1] = __vimin3_u16x2(M5[1], P1, P7);
        M3[2] = __vimax3_u16x2(P4, P5, P6); M5[2] = __vimax3_u16x2(M3[2], P3, P7); M7[2] = __vimax3_u16x2(M5[2], P2, P8);
        M3[3] = __vimin3_u16x2(P4, P5, P6); M5[3] = __vimin3_u16x2(M3[3], P3, P7); M7[3] = __vimin3_u16x2(M5[3], P2, P8);
        uint32_t out[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint32_t G, H;
          if (k & 1) {
            G = __vimin3_u16x2(m3[k][i % 2], m5[k], M7[k]);
            H = __vimin3_u16x2(m7[k][(i + 1) % 3], m5[k], M3[k]);
            out[k] = __vimin3_u16x2(gg[k][i % 4], m7[k][i % 3], H);
          } else {
            G = __vimax3_u16x2(m3[k][i % 2], m5[k], M7[k]);
            H = __vimax3_u16x2(m7[k][(i + 1) % 3], m5[k], M3[k]);
            out[k] = __vimax3_u16x2(gg[k][i % 4], m7[k][i % 3], H);
          }
          m3[k][i % 2] = M3[k]; m5[k] = M5[k]; m7[k][i % 3] = M7[k]; gg[k][i % 4] = G;
        }
        // pixels of output row y = r - 3: centre and x -+ 3 from three rows ago, (x, y - 3) from six rows ago
        const uint32_t cE = ce[(i + 3) % 6], cO = co[(i + 3) % 6], nE = ce[i % 6], nO = co[i % 6];
        const uint32_t lE = le[i % 3], lO = lo[i % 3], rE = re[i % 3], rO = ro[i % 3];
        ce[i % 6] = P4; co[i % 6] = P5; le[i % 3] = P1; lo[i % 3] = P2; re[i % 3] = P7; ro[i % 3] = P8;
        if (i < 6) continue;
        const int ry = strip * kDetStrip + i - 6;  // row inside the tile
        const int y = y0 + ry;
        if (y >= L.h) break;
        *reinterpret_cast<uint2*>(cmap + (long long)y * L.pitch + x) = make_uint2(0, 0);  // corners are written in phase 2
        const uint32_t TE = out[0] - out[1], TO = out[2] - out[3];  // per lane max >= min: no borrow
        // T of x, x+1, x+2, x+3 = TE.lo, TO.lo, TE.hi, TO.hi as four bytes
        *reinterpret_cast<uint32_t*>(&s_T[ry][cx]) = __byte_perm(TE, TO, 0x6240);
        const uint32_t bE = (uint32_t)s_b2[TE & 0xffffu] | ((uint32_t)s_b2[TE >> 16] << 16);
        const uint32_t bO = (uint32_t)s_b2[TO & 0xffffu] | ((uint32_t)s_b2[TO >> 16] << 16);
        uint32_t cand[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint32_t c = h ? cO : cE, b = h ? bO : bE;
          const uint32_t p0 = h ? lO : lE, p4 = h ? nO : nE, p8 = h ? rO : rE, p12 = h ? P5 : P4;
          // A 9-arc holds two ADJACENT compass pixels (ring positions 4 apart).  Every horizontal compass pixel is adjacent
          // to every vertical one, so the best adjacent pair is (best horizontal, best vertical): S2 = the largest value two
          // adjacent compass pixels both reach, s2 = the smallest value two adjacent ones both stay under.
          const uint32_t S2 = __vminu2(__vmaxu2(p0, p8), __vmaxu2(p4, p12)), s2 = __vmaxu2(__vminu2(p0, p8), __vminu2(p4, p12));
          // lane test A > B as ((B | 0x8000) - A) losing bit 15 (all values stay below 0x8000)
          const uint32_t bright = ((c + b) | kHi) - S2;   // bit 15 clear <=> S2 > c + b
          const uint32_t dark = ((s2 + b) | kHi) - c;     // bit 15 clear <=> c > s2 + b
          cand[h] = ~(bright & dark) & kHi;
        }
        // flags of x (bit 0), x+1 (bit 1) -> acc_lo; of x+2, x+3 -> acc_hi; the latest row in the lowest bits
        const uint32_t c4 = (cand[0] >> 15) | (cand[1] >> 14);
        acc_lo = (acc_lo << 2) | (c4 & 3u);
        acc_hi = (acc_hi << 2) | (c4 >> 16);
        last_ry = ry;
      }
      // queue the candidates of the thread's four columns (rows last_ry, last_ry - 1, ... from bit 0 upwards)
      const int n_mine = __popcll(acc_lo) + __popcll(acc_hi);
      if (n_mine) {
        int slot = atomicAdd(&s_count, n_mine);
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          unsigned long long a = half ? acc_hi : acc_lo;
          while (a) {
            const int bit = __ffsll((long long)a) - 1;
            a &= a - 1;
            const int ry = last_ry - (bit >> 1), j = 2 * half + (bit & 1);
            const int xx = x + j, y = y0 + ry;
            if (xx < 3 || xx >= L.w - 3 || y < 3 || y >= L.h - 3) continue;  // (its slot stays unused)
            if (slot < kDetQueue) s_queue[slot] = (uint16_t)((cx + j) | (ry << 8));
            else {  // queue full: test in place
              const int T = s_T[ry][cx + j];
              if (segment_test(s_img, ry + 3, sc + j, s_b2[T])) {
                cmap[(long long)y * L.pitch + xx] = (uint16_t)T;
                atomicAdd(&s_rows[ry], 1);
              }
            }
            ++slot;
          }
        }
      }
    }
  }
  __syncthreads();
  // phase 2: full segment test on the queued candidates, two per thread and step (slots of border pixels were left
  // unused: 0xffff -- such a slot is tested on a harmless interior pixel and its verdict dropped)
  const int n_cand = min(s_count, kDetQueue);
  for (int q = tid; q < n_cand; q += 2 * kDetThreads) {
    const uint32_t e0 = s_queue[q];
    const uint32_t e1 = q + kDetThreads < n_cand ? (uint32_t)s_queue[q + kDetThreads] : 0xffffu;
    const bool v0 = e0 != 0xffffu, v1 = e1 != 0xffffu;
    const int cx0 = v0 ? (int)(e0 & 0xff) : 0, ry0 = v0 ? (int)(e0 >> 8) : 0;
    const int cx1 = v1 ? (int)(e1 & 0xff) : 0, ry1 = v1 ? (int)(e1 >> 8) : 0;
    const int T0 = s_T[ry0][cx0], T1 = s_T[ry1][cx1];
    const int b0 = s_b2[T0], b1 = s_b2[T1];
    if (b0 == (int)kB2None && b1 == (int)kB2None) continue;   // (cannot happen for queued pixels; keeps the sums below 0x8000)
    const uint32_t hit = segment_test_pair(s_img, ry0 + 3, cx0 + 4, b0 == (int)kB2None ? 0x1000 : b0, ry1 + 3, cx1 + 4, b1 == (int)kB2None ? 0x1000 : b1);
    if (v0 && (hit & 1u)) {
      cmap[(long long)(y0 + ry0) * L.pitch + x0 + cx0] = (uint16_t)T0;
      atomicAdd(&s_rows[ry0], 1);
    }
    if (v1 && (hit & 2u)) {
      cmap[(long long)(y0 + ry1) * L.pitch + x0 + cx1] = (uint16_t)T1;
      atomicAdd(&s_rows[ry1], 1);
    }
  }
  __syncthreads();
  if (tid < kDetTH && s_rows[tid] && y0 + tid < L.h) atomicAdd(&rowcnt[(long long)frame * total_rows + row_off + y0 + tid], s_rows[tid]);
}

cudaError_t launch_agast_detect(const PyramidGeom& g, const DetectWorkspace& ws, int n_frames, int thresh, cudaStream_t stream,
                                int lower) {
  if (lower != kLowerThreshold && lower != 0) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemsetAsync(ws.rowcnt, 0, sizeof(int) * (size_t)n_frames * ws.total_rows, stream);
  if (e != cudaSuccess) return e;
  for (int l = 0; l < g.n_layers; ++l) {
    const LayerGeom& L = g.L[l];
    dim3 grid((L.w + kDetTW - 1) / kDetTW, (L.h + kDetTH - 1) / kDetTH, n_frames);
    if (lower == kLowerThreshold)
      agast_detect_kernel<kLowerThreshold><<<grid, kDetThreads, 0, stream>>>(L, g.frame_elems, ws.pyr, ws.cm, ws.rowcnt, ws.total_rows, ws.row_off[l], thresh);
    else
      agast_detect_kernel<0><<<grid, kDetThreads, 0, stream>>>(L, g.frame_elems, ws.pyr, ws.cm, ws.rowcnt, ws.total_rows, ws.row_off[l], thresh);
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
  for (int xb = 0; xb < L.w; xb += 256) {
    // eight map entries (16 bytes) per lane; rows are sparse, so most steps end at the first ballot
    const int x = xb + 8 * lane;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (x < L.pitch) v = *reinterpret_cast<const uint4*>(row + x);
    if (!__ballot_sync(0xffffffffu, (v.x | v.y | v.z | v.w) != 0)) continue;   // 256 empty entries: the common case
    const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
    uint32_t m8 = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if ((wv[j] & 0xffffu) && x + 2 * j < L.w) m8 |= 1u << (2 * j);
      if ((wv[j] >> 16) && x + 2 * j + 1 < L.w) m8 |= 2u << (2 * j);
    }
    if (!__ballot_sync(0xffffffffu, m8 != 0)) continue;
    const int cnt = __popc(m8);
    int inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
    int s = slot + inc - cnt;
    for (uint32_t m = m8; m; m &= m - 1) {
      const int j = __ffs(m) - 1;
      if (s < corner_cap) out[s] = (uint32_t)(x + j) | ((uint32_t)y << 13) | ((uint32_t)layer << 26);
      ++s;
    }
    slot += __shfl_sync(0xffffffffu, inc, 31);
  }
}

// Guard for AGAST thresholds below 20: the closed form of the lazy score cache (nms_logic.cuh) needs every detected
// corner to hold a score > 2 (only such cache entries are sticky, brisk-layer.cc:124-126).  For thresh >= 20 that is
// guaranteed; below, it is checked on the actual corners: *flag = 5 when one of them scores <= 2.
__global__ void __launch_bounds__(256)
corner_score_check_kernel(PyramidGeom g, DetectWorkspace ws, int* __restrict__ flag) {
  const int frame = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = min(ws.layer_start[(long long)frame * (kMaxLayers + 1) + g.n_layers], ws.corner_cap);
  if (k >= n) return;
  const uint32_t c = ws.corners[(long long)frame * ws.corner_cap + k];
  const int x = c & 0x1fff, y = (c >> 13) & 0x1fff, layer = c >> 26;
  const LayerGeom& L = g.L[layer];
  const int T = ws.cm[(long long)frame * g.frame_elems + L.off + (long long)y * L.pitch + x] & kCmT;
  if (T <= 2) atomicExch(flag, 5);
}

cudaError_t launch_corner_score_check(const PyramidGeom& g, const DetectWorkspace& ws, int n_frames, int* flag, cudaStream_t stream) {
  dim3 grid((ws.corner_cap + 255) / 256, n_frames);
  corner_score_check_kernel<<<grid, 256, 0, stream>>>(g, ws, flag);
  return cudaGetLastError();
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
