// Tensor-core variant of the brute-force Hamming 2-NN search (sm_100a), kept next to the POPC
// kernel of hamming.cu so the two can be measured against each other (BASELINE.json: "a measured
// int8 tensor-core GEMM variant kept only if ncu shows it wins").
//
// Blackwell has no binary MMA (`mma.sync ... b1 ... and.popc` is expanded by ptxas into eight
// IMMA.16832.U8), so the descriptors' bits are expanded to bytes while they are staged in shared
// memory and the kernel runs an s8 x u8 -> s32 GEMM with `mma.sync.m16n8k32`: query bits become
// +1/-1, train bits 0/1, so that
//   v(q, t) = sum_k t_k (2 q_k - 1) = 2 popcount(q AND t) - popcount(t),  hamming(q, t) = popcount(q) - v(q, t).
// A CTA owns 128 queries (expanded once) and streams 128-row train tiles through a double buffer
// filled by four producer warps; each of the eight consumer warps computes a 64 x 32 block of v,
// and keeps the two smallest (distance, train index) pairs per query and thread in registers.  Keys
// order exactly like the reference's selection rule (first minimum wins => lowest train index on ties).
#include <cuda_runtime.h>

#include "kernels.h"

namespace briskb200 {

constexpr int kMmaTile = 128;        // queries per CTA, train rows per tile
constexpr int kMmaConsumers = 256;   // 8 MMA warps: 2 (M) x 4 (N), warp tile 64 x 32
constexpr int kMmaProducers = 128;   // 4 staging warps: one train row per thread and tile
constexpr int kMmaThreads = kMmaConsumers + kMmaProducers;
constexpr int kMmaPad = 16;          // row padding (bytes) -> conflict-free ldmatrix rows and 128-bit stores
constexpr unsigned long long kMmaKeyNone = ~0ull;
constexpr unsigned kMmaDistNone = 0x7fff0000u;   // list sentinel; every real distance is <= 8 * DB
constexpr int kBarFull = 1, kBarEmpty = 3;       // named barriers: kBarFull + buf, kBarEmpty + buf

__device__ __forceinline__ void bar_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kMmaThreads) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(kMmaThreads) : "memory"); }

// 4 descriptor bits -> 4 bytes of 0/1 (byte i = bit i of the low nibble of b).
// The multiply parks bit i in the top bit of byte i; PRMT's sign-replicate mode turns that into 0x00 / 0xff.
template <bool SIGNED>
__device__ __forceinline__ unsigned expand_nibble(unsigned b) {
  unsigned mask;   // (__byte_perm drops the selector's replicate-sign bits, hence the PTX)
  asm("prmt.b32 %0, %1, %1, 0xba98;" : "=r"(mask) : "r"((b & 0xfu) * 0x10204080u));
  return SIGNED ? (mask & 0x01010101u) | ~mask   // bit set -> 0x01, clear -> 0xff (-1)
                : mask & 0x01010101u;
}

// 32 descriptor bits -> 32 bytes (0/1, or +1/-1 when SIGNED), stored as two 128-bit words.
template <bool SIGNED>
__device__ __forceinline__ void store_expanded_word(uint8_t* dst, unsigned v) {
  uint4 lo, hi;
  lo.x = expand_nibble<SIGNED>(v); lo.y = expand_nibble<SIGNED>(v >> 4); lo.z = expand_nibble<SIGNED>(v >> 8); lo.w = expand_nibble<SIGNED>(v >> 12);
  hi.x = expand_nibble<SIGNED>(v >> 16); hi.y = expand_nibble<SIGNED>(v >> 20); hi.z = expand_nibble<SIGNED>(v >> 24); hi.w = expand_nibble<SIGNED>(v >> 28);
  *reinterpret_cast<uint4*>(dst) = lo;
  *reinterpret_cast<uint4*>(dst + 16) = hi;
}

template <int DB>
struct RowRegs { uint4 v[DB / 16]; };

template <int DB>
__device__ __forceinline__ RowRegs<DB> load_row(const uint8_t* __restrict__ src, bool valid) {
  RowRegs<DB> r;
#pragma unroll
  for (int i = 0; i < DB / 16; ++i) r.v[i] = valid ? __ldg(reinterpret_cast<const uint4*>(src) + i) : make_uint4(0, 0, 0, 0);
  return r;
}

// Expand one descriptor row into dst (DB * 8 bytes); returns its popcount.
template <int DB, bool SIGNED>
__device__ __forceinline__ int store_row(uint8_t* dst, const RowRegs<DB>& r) {
  int pc = 0;
#pragma unroll
  for (int i = 0; i < DB / 16; ++i) {
    const unsigned w[4] = {r.v[i].x, r.v[i].y, r.v[i].z, r.v[i].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      store_expanded_word<SIGNED>(dst + (i * 4 + j) * 32, w[j]);
      pc += __popc(w[j]);
    }
  }
  return pc;
}

__device__ __forceinline__ void ldmatrix_x4(unsigned addr, unsigned& r0, unsigned& r1, unsigned& r2, unsigned& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}

// Shared memory: sA [128][LD] expanded queries | sB [2][128][LD] expanded train tiles (double buffered; the
// first 32 KB are reused for the final list merge) | pq [128].
template <int DB>
__global__ void __launch_bounds__(kMmaThreads, 1)
hamming_knn2_mma_kernel(const uint8_t* __restrict__ q, long long nq, const uint8_t* __restrict__ t, long long nt,
                        long long rows_per_split, long long train_index_offset, unsigned long long* __restrict__ part) {
  constexpr int LD = DB * 8 + kMmaPad;
  extern __shared__ __align__(16) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = sA + kMmaTile * LD;
  int* pq = reinterpret_cast<int*>(sB + 2 * kMmaTile * LD);

  const int tid = threadIdx.x;
  const long long q0 = (long long)blockIdx.x * kMmaTile;
  const long long t_begin = (long long)blockIdx.y * rows_per_split;
  const long long t_end = min(nt, t_begin + rows_per_split);
  const int ntiles = t_end > t_begin ? (int)((t_end - t_begin + kMmaTile - 1) / kMmaTile) : 0;

  if (tid < kMmaTile) {
    const bool valid = q0 + tid < nq;
    const RowRegs<DB> r = load_row<DB>(q + (q0 + tid) * DB, valid);
    pq[tid] = store_row<DB, true>(sA + tid * LD, r);
  }
  __syncthreads();

  unsigned d0[8], d1[8], i0[8], i1[8];   // consumers: two best (distance, train index) per owned row
#pragma unroll
  for (int i = 0; i < 8; ++i) { d0[i] = kMmaDistNone; d1[i] = kMmaDistNone; i0[i] = 0xffffffffu; i1[i] = 0xffffffffu; }

  if (tid >= kMmaConsumers) {
    // ---- producers: thread p stages train row p of every tile ----
    const int p = tid - kMmaConsumers;
    RowRegs<DB> cur = load_row<DB>(t + (t_begin + p) * DB, ntiles > 0 && t_begin + p < t_end);
    for (int i = 0; i < ntiles; ++i) {
      const int buf = i & 1;
      const long long next_row = t_begin + (long long)(i + 1) * kMmaTile + p;
      const RowRegs<DB> nxt = load_row<DB>(t + next_row * DB, next_row < t_end);
      if (i >= 2) bar_sync(kBarEmpty + buf);
      store_row<DB, false>(sB + (buf * kMmaTile + p) * LD, cur);
      bar_arrive(kBarFull + buf);
      cur = nxt;
    }
  } else {
    // ---- consumers ----
    const int lane = tid & 31, warp = tid >> 5;
    const int tq = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;
    const unsigned sA_u = (unsigned)__cvta_generic_to_shared(sA);
    const unsigned sB_u = (unsigned)__cvta_generic_to_shared(sB);
    // ldmatrix row addresses: A fragment (16 rows x 32 k-bytes): matrices (rows 0-7,k 0-15) (rows 8-15,k 0-15) (rows 0-7,k 16-31) (rows 8-15,k 16-31)
    const unsigned a_off = (unsigned)((wm * 64 + (lane & 7) + 8 * ((lane >> 3) & 1)) * LD + 16 * (lane >> 4));
    // B fragments of two 8-column blocks: (cols 0-7,k 0-15) (cols 0-7,k 16-31) (cols 8-15,k 0-15) (cols 8-15,k 16-31)
    const unsigned b_off = (unsigned)((wn * 32 + (lane & 7) + 8 * (lane >> 4)) * LD + 16 * ((lane >> 3) & 1));
    const int g = lane >> 2;
    int pqr[8], vthr[8];   // popcount of the row's query; v must exceed vthr = pq - (2nd best distance) to matter
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) { pqr[mt * 2] = pq[wm * 64 + mt * 16 + g]; pqr[mt * 2 + 1] = pq[wm * 64 + mt * 16 + g + 8]; }
#pragma unroll
    for (int i = 0; i < 8; ++i) vthr[i] = pqr[i] - (int)kMmaDistNone;

    for (int i = 0; i < ntiles; ++i) {
      const int buf = i & 1;
      bar_sync(kBarFull + buf);
      int acc[4][4][4];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int n = 0; n < 4; ++n) { acc[mt][n][0] = 0; acc[mt][n][1] = 0; acc[mt][n][2] = 0; acc[mt][n][3] = 0; }
      const unsigned a_base = sA_u + a_off, b_base = sB_u + (unsigned)(buf * kMmaTile * LD) + b_off;
#pragma unroll 2
      for (int k0 = 0; k0 < DB * 8; k0 += 32) {
        unsigned a[4][4], b[4][2];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) ldmatrix_x4(a_base + mt * 16 * LD + k0, a[mt][0], a[mt][1], a[mt][2], a[mt][3]);
#pragma unroll
        for (int np = 0; np < 2; ++np) ldmatrix_x4(b_base + np * 16 * LD + k0, b[2 * np][0], b[2 * np][1], b[2 * np + 1][0], b[2 * np + 1][1]);
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
          for (int n = 0; n < 4; ++n)
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+r"(acc[mt][n][0]), "+r"(acc[mt][n][1]), "+r"(acc[mt][n][2]), "+r"(acc[mt][n][3])
                         : "r"(a[mt][0]), "r"(a[mt][1]), "r"(a[mt][2]), "r"(a[mt][3]), "r"(b[n][0]), "r"(b[n][1]));
      }
      if (i + 2 < ntiles) bar_arrive(kBarEmpty + buf);
      // epilogue: one max + compare per row decides whether any of its 8 columns can enter the list.
      // A thread meets the train rows of one query in increasing index order, so an equal distance
      // never displaces an earlier entry.
      const long long tile_base = t_begin + (long long)i * kMmaTile;
      const int valid_cols = (int)min((long long)kMmaTile, t_end - tile_base);
      const unsigned col_base = (unsigned)(train_index_offset + tile_base);
      const int col_thread = wn * 32 + 2 * tq;
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int ri = mt * 2 + half;
          const int h = half * 2;
          const int m = max(max(max(acc[mt][0][h], acc[mt][0][h + 1]), max(acc[mt][1][h], acc[mt][1][h + 1])),
                            max(max(acc[mt][2][h], acc[mt][2][h + 1]), max(acc[mt][3][h], acc[mt][3][h + 1])));
          if (m > vthr[ri]) {
#pragma unroll
            for (int n = 0; n < 4; ++n)
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                const int col = col_thread + n * 8 + j;
                const unsigned ham = (unsigned)(pqr[ri] - acc[mt][n][h + j]);
                if (ham < d1[ri] && col < valid_cols) {
                  const unsigned idx = col_base + col;
                  if (ham < d0[ri]) { d1[ri] = d0[ri]; i1[ri] = i0[ri]; d0[ri] = ham; i0[ri] = idx; }
                  else { d1[ri] = ham; i1[ri] = idx; }
                }
              }
            vthr[ri] = pqr[ri] - (int)d1[ri];
          }
        }
    }
  }
  __syncthreads();
  // merge the 16 owner lists of every row (through the now idle train buffers)
  unsigned long long* lists = reinterpret_cast<unsigned long long*>(sB);   // [128 rows][16 owners][2]
  if (tid < kMmaConsumers) {
    const int lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tq = lane & 3, wm = warp >> 2, wn = warp & 3;
    const int owner = wn * 4 + tq;
#pragma unroll
    for (int ri = 0; ri < 8; ++ri) {
      const int row = wm * 64 + (ri >> 1) * 16 + g + 8 * (ri & 1);
      unsigned long long* l = lists + (row * 16 + owner) * 2;
      l[0] = d0[ri] == kMmaDistNone ? kMmaKeyNone : ((unsigned long long)d0[ri] << 32) | i0[ri];
      l[1] = d1[ri] == kMmaDistNone ? kMmaKeyNone : ((unsigned long long)d1[ri] << 32) | i1[ri];
    }
  }
  __syncthreads();
  unsigned long long* out = part + (long long)blockIdx.y * nq * 2;
  if (tid < kMmaTile && q0 + tid < nq) {
    unsigned long long b0 = kMmaKeyNone, b1 = kMmaKeyNone;
    for (int o = 0; o < 32; ++o) {
      const unsigned long long key = lists[tid * 32 + ((o + tid) & 31)];
      if (key < b0) { b1 = b0; b0 = key; } else if (key < b1) { b1 = key; }
    }
    out[(q0 + tid) * 2] = b0; out[(q0 + tid) * 2 + 1] = b1;
  }
}

template <int DB>
static cudaError_t launch_mma(const uint8_t* q, long long nq, const uint8_t* t, long long nt, long long off,
                              unsigned long long* dst, int splits, long long rows_per_split, cudaStream_t stream) {
  constexpr int LD = DB * 8 + kMmaPad;
  const size_t smem = (size_t)3 * kMmaTile * LD + kMmaTile * sizeof(int);
  cudaError_t e = cudaFuncSetAttribute(hamming_knn2_mma_kernel<DB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  dim3 grid((unsigned)((nq + kMmaTile - 1) / kMmaTile), splits);
  hamming_knn2_mma_kernel<DB><<<grid, kMmaThreads, smem, stream>>>(q, nq, t, nt, rows_per_split, off, dst);
  return cudaGetLastError();
}

// Number of train-set splits: enough CTAs to fill the 148 SMs, then the smallest count whose last wave is
// at least 95% full (one CTA per SM; every split costs one more pass over the query tile and a merge row).
int knn_mma_num_splits(long long nq, long long nt) {
  const long long qblocks = (nq + kMmaTile - 1) / kMmaTile;
  long long max_splits = (nt + 8 * kMmaTile - 1) / (8 * kMmaTile);
  if (max_splits > 64) max_splits = 64;
  if (max_splits < 1) max_splits = 1;
  long long best = 1;
  double best_eff = 0.0;
  for (long long s = 1; s <= max_splits; ++s) {
    const long long ctas = qblocks * s;
    const long long waves = (ctas + 147) / 148;
    const double eff = (double)ctas / (double)(waves * 148);
    if (eff > best_eff + 1e-9) { best_eff = eff; best = s; }
    if (eff >= 0.95) break;
  }
  return (int)best;
}

// k == 2 only; keys [nq][2]; part: scratch [splits][nq][2] (unused when splits == 1).
cudaError_t launch_hamming_knn2_mma(const uint8_t* q, long long nq, const uint8_t* t, long long nt, int desc_bytes,
                                    long long train_index_offset, unsigned long long* keys, unsigned long long* part,
                                    int splits, cudaStream_t stream) {
  if (nq <= 0) return cudaSuccess;
  long long rows_per_split = ((nt + splits - 1) / splits + kMmaTile - 1) / kMmaTile * kMmaTile;
  if (rows_per_split <= 0) rows_per_split = kMmaTile;
  unsigned long long* dst = splits == 1 ? keys : part;
  cudaError_t e;
  if (desc_bytes == 64) e = launch_mma<64>(q, nq, t, nt, train_index_offset, dst, splits, rows_per_split, stream);
  else if (desc_bytes == 48) e = launch_mma<48>(q, nq, t, nt, train_index_offset, dst, splits, rows_per_split, stream);
  else return cudaErrorInvalidValue;
  if (e != cudaSuccess) return e;
  if (splits > 1) e = launch_knn_merge(part, splits, nq, 2, keys, stream);
  return e;
}

}  // namespace briskb200
