// Brute-force Hamming 2-NN on the 5th-generation tensor cores (sm_100a): tcgen05.mma kind::i8 with TMEM
// accumulators, operands staged by TMA, top-2 selection straight out of TMEM.
//
// Replaces the inner loops of BruteForceMatcher::commonKnnMatchImpl (reference brisk/src/brute-force-matcher.cc:
// 80-162) and Hamming::SSSE3PopcntofXORed (brisk/include/brisk/internal/hamming-inl.h:85-134) for k = 2 and
// 48 / 64-byte rows, like hamming_mma.cu (mma.sync IMMA) and hamming.cu (XOR + POPC), which stay as the measured
// alternatives.
//
// Formulation.  Every descriptor bit becomes one signed byte, +1 for a set bit and -1 for a clear one (expanded ONCE
// per call into HBM, 8 bytes per descriptor byte).  For two rows q, t of K bits,  q . t = K - 2 hamming(q, t),  so the
// nearest neighbours are the largest dot products and  hamming = (K - q . t) / 2  exactly.
//
// Kernel.  A CTA owns 256 queries (two 128-row A tiles, resident in shared memory for the whole kernel: 128 KB for
// K = 512) and streams its share of the train set in tiles of 128 rows, cut along K into 128-byte chunks (one
// 16 KB TMA box each, 128B swizzle, K-major) through a six-stage ring.  One thread issues the MMAs: per chunk four
// K = 32 steps for each A tile, 128 x 128 x 32 each, accumulating in TMEM; the two accumulator sets (2 x 256 columns =
// all 512 TMEM columns) let the epilogue of tile i overlap the MMAs of tile i + 1.  Eight epilogue warps (one per
// TMEM lane quadrant and A tile) read the int32 dot products with tcgen05.ld; a thread owns ONE query for the whole
// kernel, keeps its two best (dot, train index) pairs in registers and only looks closer at a group of 32 columns
// when the group's maximum beats its current second best (a few dozen times per query over millions of rows).
// Train rows are met in increasing index order, so a strict comparison reproduces the reference's tie rule (first
// minimum of a left-to-right scan wins).  Per-split results go to `part` as (hamming << 32 | index) keys, merged by
// launch_knn_merge like the other variants.
//
// Roofline: 512 int8 MACs per 512-bit comparison on the tensor pipe; B traffic from L2 is
// 128 x 512 B per 256 x 128 comparisons (M = 256 per CTA halves what a single 128-row tile would pull).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "brisk_math.cuh"
#include "kernels.h"

namespace briskb200 {

namespace {

constexpr int kT5M = 128;         // rows of one MMA (= TMEM lanes)
constexpr int kT5QTiles = 2;      // A tiles per CTA -> 256 queries
// Train rows per tile (MMA N): 128, or 256 with BRISK_B200_TC5_TILE_ROWS=256 in the environment (kept for measurements).
static int t5_tile_rows() {
  static const int rows = [] { const char* e = getenv("BRISK_B200_TC5_TILE_ROWS"); return e && atoi(e) == 256 ? 256 : 128; }();
  return rows;
}
constexpr int kT5Chunk = 128;     // bytes of K per shared-memory chunk: one row of a 128B swizzle atom
constexpr int kT5AChunkBytes = kT5M * kT5Chunk;  // 16 KB
constexpr int kT5RingBytes = 96 * 1024;          // B ring: 6 stages of 128 train rows or 3 stages of 256
constexpr int kT5EpiWarps = 8;
constexpr int kT5Threads = (2 + kT5EpiWarps) * 32;  // warp 0: TMA producer, warp 1: TMEM owner + MMA issuer, 2..9: epilogue
constexpr int kT5TmemCols = 512;
constexpr unsigned long long kT5KeyNone = ~0ull;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Shared-memory matrix descriptor of a K-major operand chunk stored as rows of 128 bytes under the 128B swizzle
// (what a TMA box of 128 bytes x R rows with CU_TENSOR_MAP_SWIZZLE_128B writes): start address >> 4 in bits 0-13,
// leading byte offset (unused for swizzled K-major layouts) 1, stride byte offset = 1024 B between 8-row groups in
// bits 32-45, descriptor version 1 (sm_100) in bits 46-47, layout type 2 = SWIZZLE_128B in bits 61-63.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3ffffu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// Instruction descriptor, kind::i8: D = S32 (2 << 4), A = B = signed 8 bit (1 << 7, 1 << 10), both K-major, N >> 3 in bits
// 17-22, M >> 4 in bits 24-28.
__host__ __device__ constexpr uint32_t t5_idesc(int n) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kT5M >> 4) << 24);
}

__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}
// All MMAs issued so far by this thread arrive on `bar` when they complete (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 consecutive TMEM columns of the warp's 32 lanes: thread = lane (row), v[j] = column j.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, int v[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// The same load without the wait (several in flight), and the wait for all of a thread's outstanding loads.
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, int v[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ int max3(int a, int b, int c) { return max(max(a, b), c); }

}  // namespace

// KC = chunks along K: 4 for 64-byte rows (512 signed bytes), 3 for 48-byte rows.
// NT = train rows per tile = N of one MMA.  128 (default): two accumulator sets, the epilogue of tile i runs under the MMAs of
// tile i + 1.  Every 64-cycle MMA reads 4 KB of A and 4 KB of B from shared memory while TMA writes 2 KB of the next chunk --
// more than the 128 B / cycle an SM's shared memory delivers, which is what bounds the kernel (measured: tensor pipe 68 %
// active, 2.75 Tcmp/s at 512 bit, 4.2 Tcmp/s at 384 bit).  256: the same bytes per MAC, and the accumulators of the two A
// tiles fill all 512 TMEM columns, so the next tile's MMAs wait for the drain (measured: 2.53 / 3.04 Tcmp/s).
template <int KC, int NT>
__global__ void __launch_bounds__(kT5Threads, 1)
hamming_knn2_tc5_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_t, long long nq,
                        long long nt, long long rows_per_split, long long train_index_offset,
                        unsigned long long* __restrict__ part) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);  // swizzle atoms are 1024-byte aligned
  constexpr int kT5N = NT, kT5ChunkBytes = NT * kT5Chunk, kT5Stages = kT5RingBytes / kT5ChunkBytes;
  constexpr int kSets = kT5TmemCols / (kT5QTiles * NT);       // accumulator sets in TMEM: 2 (NT = 128) or 1 (NT = 256)
  constexpr uint32_t kIdesc = t5_idesc(NT);
  uint8_t* sA = smem;                                         // [2 tiles][KC chunks][128 rows x 128 B]
  uint8_t* sB = sA + kT5QTiles * KC * kT5AChunkBytes;         // [stages][NT rows x 128 B]
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(sB + kT5Stages * kT5ChunkBytes);
  uint64_t* bar_empty = bar_full + kT5Stages;
  uint64_t* bar_a = bar_empty + kT5Stages;
  uint64_t* bar_tfull = bar_a + 1;    // [2] accumulator set ready for the epilogue
  uint64_t* bar_tempty = bar_tfull + 2;  // [2] accumulator set drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tempty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long q0 = (long long)blockIdx.x * (kT5QTiles * kT5M);
  const long long t_begin = (long long)blockIdx.y * rows_per_split;
  const long long t_end = min(nt, t_begin + rows_per_split);
  const int ntiles = t_end > t_begin ? (int)((t_end - t_begin + kT5N - 1) / kT5N) : 0;

  if (tid == 0) {
    for (int s = 0; s < kT5Stages; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
    mbar_init(bar_a, 1);
    for (int b = 0; b < 2; ++b) { mbar_init(&bar_tfull[b], 1); mbar_init(&bar_tempty[b], kT5EpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kT5TmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---- TMA producer ----
    if (lane == 0 && ntiles > 0) {
      mbar_expect_tx(bar_a, kT5QTiles * KC * kT5AChunkBytes);
      for (int a = 0; a < kT5QTiles; ++a)
        for (int c = 0; c < KC; ++c) tma_load_2d(sA + (a * KC + c) * kT5AChunkBytes, &map_q, c * kT5Chunk, (int)(q0 + a * kT5M), bar_a);
      int it = 0;
      for (int i = 0; i < ntiles; ++i)
        for (int c = 0; c < KC; ++c, ++it) {
          const int s = it % kT5Stages;
          mbar_wait(&bar_empty[s], ((it / kT5Stages) & 1) ^ 1);   // a fresh barrier passes the parity-1 wait
          mbar_expect_tx(&bar_full[s], kT5ChunkBytes);
          tma_load_2d(sB + s * kT5ChunkBytes, &map_t, c * kT5Chunk, (int)(t_begin + (long long)i * kT5N), &bar_full[s]);
        }
    }
  } else if (warp == 1) {
    // ---- MMA issuer (one thread) ----
    if (lane == 0 && ntiles > 0) {
      mbar_wait(bar_a, 0);
      tc_fence_after();
      const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
      int it = 0;
      for (int i = 0; i < ntiles; ++i) {
        const int b = i % kSets;
        mbar_wait(&bar_tempty[b], ((i / kSets) & 1) ^ 1);
        tc_fence_after();
        for (int c = 0; c < KC; ++c, ++it) {
          const int s = it % kT5Stages;
          mbar_wait(&bar_full[s], (it / kT5Stages) & 1);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < kT5Chunk / 32; ++k) {
            const uint64_t bd = umma_desc(b_base + s * kT5ChunkBytes + k * 32);
#pragma unroll
            for (int a = 0; a < kT5QTiles; ++a)
              umma_i8(tmem_base + (uint32_t)((b * kT5QTiles + a) * kT5N), umma_desc(a_base + (a * KC + c) * kT5AChunkBytes + k * 32), bd,
                      kIdesc, (c | k) != 0 ? 1u : 0u);
          }
          umma_commit(&bar_empty[s]);   // the stage is free once these MMAs have read it
        }
        umma_commit(&bar_tfull[b]);     // both accumulators of the set are complete
      }
    }
  } else {
    // ---- epilogue: thread = one query row ----
    const int quad = warp & 3, a = (warp - 2) >> 2;
    const int row = a * kT5M + quad * 32 + lane;
    int d0 = -100000, d1 = -100000;       // two largest dot products so far (d0 >= d1) ...
    unsigned i0 = 0xffffffffu, i1 = 0xffffffffu;  // ... and their (global) train indices
    for (int i = 0; i < ntiles; ++i) {
      const int b = i % kSets;
      mbar_wait(&bar_tfull[b], (i / kSets) & 1);
      tc_fence_after();
      const long long tile_base = t_begin + (long long)i * kT5N;
      const int valid = (int)min((long long)kT5N, t_end - tile_base);
      const unsigned idx_base = (unsigned)(train_index_offset + tile_base);
#pragma unroll 1
      for (int cc = 0; cc < kT5N / 32; ++cc) {
        int v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)((b * kT5QTiles + a) * kT5N + cc * 32), v);
        int m = max3(v[0], v[1], v[2]);
#pragma unroll
        for (int j = 3; j + 1 < 32; j += 2) m = max3(m, v[j], v[j + 1]);
        m = max(m, v[31]);
        if (m > d1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = cc * 32 + j;
            if (v[j] > d1 && col < valid) {
              if (v[j] > d0) { d1 = d0; i1 = i0; d0 = v[j]; i0 = idx_base + col; }
              else { d1 = v[j]; i1 = idx_base + col; }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_tempty[b]);
    }
    if (q0 + row < nq) {
      constexpr int kBits = KC * kT5Chunk;
      unsigned long long* out = part + ((long long)blockIdx.y * nq + q0 + row) * 2;
      out[0] = i0 == 0xffffffffu ? kT5KeyNone : ((unsigned long long)((kBits - d0) >> 1) << 32) | i0;
      out[1] = i1 == 0xffffffffu ? kT5KeyNone : ((unsigned long long)((kBits - d1) >> 1) << 32) | i1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kT5TmemCols) : "memory");
  }
}

// ---------------------------------------------------------------------------
// A-in-TMEM form (tcgen05.mma with the A operand read from tensor memory).  The shared-memory form above is bound by the
// SM's shared-memory bandwidth: every MMA re-reads its 4 KB A slice.  Here the two query tiles are written into TMEM once,
// by the epilogue threads themselves (measured alternative, see knn_tc5_queries_in_tmem; thread = query row: it expands its descriptor's bits to +-1 bytes in registers and
// stores them with tcgen05.st; no expanded copy of the queries in HBM, no A tiles in shared memory), and shared memory
// only streams B: 4 KB read per 64-cycle MMA plus the TMA refill, 96 B / cycle.  TMEM columns: A tile 0 at 0, A tile 1 at
// 128 (K / 4 columns each, 4 signed bytes per column), accumulator of tile 0 at 256, of tile 1 at 384 -- ONE accumulator
// per A tile, yet no MMA ever waits for the epilogue: per train tile the issuer runs all K steps of A tile 0, then all of
// A tile 1 (the B tile stays in the ring for both), so each accumulator is drained while the other one is being computed.
// The freed shared memory holds a 13-stage ring (three train tiles in flight).
// ---------------------------------------------------------------------------
constexpr int kTsStages = 13;

__device__ __forceinline__ void umma_i8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t v[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
        "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
        "r"(v[30]), "r"(v[31])
      : "memory");
}

// 4 descriptor bits -> 4 signed bytes (+1 / -1), byte i = bit i
__device__ __forceinline__ uint32_t pm1_nibble(uint32_t nib) {
  uint32_t mask;
  asm("prmt.b32 %0, %1, %1, 0xba98;" : "=r"(mask) : "r"((nib & 0xfu) * 0x10204080u));
  return (mask & 0x01010101u) | ~mask;
}

template <int KC>
__global__ void __launch_bounds__(kT5Threads, 1)
hamming_knn2_tc5ts_kernel(const uint8_t* __restrict__ q, const __grid_constant__ CUtensorMap map_t, long long nq, long long nt,
                          long long rows_per_split, long long train_index_offset, unsigned long long* __restrict__ part) {
  constexpr int kN = 128, kChunkBytes = kN * kT5Chunk, kDB = KC * 16;   // descriptor bytes per row
  constexpr uint32_t kIdesc = t5_idesc(kN);
  constexpr uint32_t kColA = 0, kColD = 256;   // A tile a at kColA + 128 a, accumulator a at kColD + 128 a
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = smem;                                          // [stages][128 rows x 128 B]
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(sB + kTsStages * kChunkBytes);
  uint64_t* bar_empty = bar_full + kTsStages;
  uint64_t* bar_a = bar_empty + kTsStages;      // both A tiles are in TMEM
  uint64_t* bar_tfull = bar_a + 1;              // [2] accumulator of A tile a complete
  uint64_t* bar_tempty = bar_tfull + 2;         // [2] accumulator of A tile a drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tempty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long q0 = (long long)blockIdx.x * (kT5QTiles * kT5M);
  const long long t_begin = (long long)blockIdx.y * rows_per_split;
  const long long t_end = min(nt, t_begin + rows_per_split);
  const int ntiles = t_end > t_begin ? (int)((t_end - t_begin + kN - 1) / kN) : 0;

  if (tid == 0) {
    for (int s = 0; s < kTsStages; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
    mbar_init(bar_a, kT5EpiWarps);
    for (int a = 0; a < 2; ++a) { mbar_init(&bar_tfull[a], 1); mbar_init(&bar_tempty[a], kT5EpiWarps / 2); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kT5TmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---- TMA producer: chunk c of tile i goes to ring slot (i * KC + c) % stages ----
    if (lane == 0) {
      int it = 0;
      for (int i = 0; i < ntiles; ++i)
        for (int c = 0; c < KC; ++c, ++it) {
          const int s = it % kTsStages;
          mbar_wait(&bar_empty[s], ((it / kTsStages) & 1) ^ 1);
          mbar_expect_tx(&bar_full[s], kChunkBytes);
          tma_load_2d(sB + s * kChunkBytes, &map_t, c * kT5Chunk, (int)(t_begin + (long long)i * kN), &bar_full[s]);
        }
    }
  } else if (warp == 1) {
    // ---- MMA issuer ----
    if (lane == 0 && ntiles > 0) {
      mbar_wait(bar_a, 0);
      tc_fence_after();
      const uint32_t b_base = smem_u32(sB);
      for (int i = 0; i < ntiles; ++i) {
#pragma unroll 1
        for (int a = 0; a < kT5QTiles; ++a) {
          mbar_wait(&bar_tempty[a], (i & 1) ^ 1);
          tc_fence_after();
          for (int c = 0; c < KC; ++c) {
            const int it = i * KC + c, s = it % kTsStages;
            if (a == 0) { mbar_wait(&bar_full[s], (it / kTsStages) & 1); tc_fence_after(); }
#pragma unroll
            for (int k = 0; k < kT5Chunk / 32; ++k)
              umma_i8_ts(tmem_base + kColD + (uint32_t)(a * kN), tmem_base + kColA + (uint32_t)(a * 128 + (c * 4 + k) * 8),
                         umma_desc(b_base + s * kChunkBytes + k * 32), kIdesc, (c | k) != 0 ? 1u : 0u);
            if (a == kT5QTiles - 1) umma_commit(&bar_empty[s]);   // both A tiles have read the chunk
          }
          umma_commit(&bar_tfull[a]);
        }
      }
    }
  } else {
    // ---- epilogue warps: thread = one query row; first the row goes into TMEM, then the top-2 selection ----
    const int quad = warp & 3, a = (warp - 2) >> 2;
    const int row = a * kT5M + quad * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    {
      const bool valid = q0 + row < nq;
      const uint4* src = reinterpret_cast<const uint4*>(q + (q0 + row) * kDB);
#pragma unroll 1
      for (int g = 0; g < kDB / 16; ++g) {   // 16 descriptor bytes = 128 bits -> 128 signed bytes = 32 TMEM columns
        const uint4 bits = valid ? __ldg(src + g) : make_uint4(0, 0, 0, 0);
        const uint32_t wds[4] = {bits.x, bits.y, bits.z, bits.w};
        uint32_t v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = valid ? pm1_nibble(wds[j >> 3] >> (4 * (j & 7))) : 0u;
        tmem_st32(lane_addr + kColA + (uint32_t)(a * 128 + g * 32), v);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_a);
    }
    int d0 = -100000, d1 = -100000;
    unsigned i0 = 0xffffffffu, i1 = 0xffffffffu;
    for (int i = 0; i < ntiles; ++i) {
      mbar_wait(&bar_tfull[a], i & 1);
      tc_fence_after();
      const long long tile_base = t_begin + (long long)i * kN;
      const int valid = (int)min((long long)kN, t_end - tile_base);
      const unsigned idx_base = (unsigned)(train_index_offset + tile_base);
#pragma unroll 1
      for (int cc = 0; cc < kN / 32; ++cc) {
        int v[32];
        tmem_ld32(lane_addr + kColD + (uint32_t)(a * kN + cc * 32), v);
        int m = max3(v[0], v[1], v[2]);
#pragma unroll
        for (int j = 3; j + 1 < 32; j += 2) m = max3(m, v[j], v[j + 1]);
        m = max(m, v[31]);
        if (m > d1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = cc * 32 + j;
            if (v[j] > d1 && col < valid) {
              if (v[j] > d0) { d1 = d0; i1 = i0; d0 = v[j]; i0 = idx_base + col; }
              else { d1 = v[j]; i1 = idx_base + col; }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_tempty[a]);
    }
    if (q0 + row < nq) {
      constexpr int kBits = KC * kT5Chunk;
      unsigned long long* out = part + ((long long)blockIdx.y * nq + q0 + row) * 2;
      out[0] = i0 == 0xffffffffu ? kT5KeyNone : ((unsigned long long)((kBits - d0) >> 1) << 32) | i0;
      out[1] = i1 == 0xffffffffu ? kT5KeyNone : ((unsigned long long)((kBits - d1) >> 1) << 32) | i1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kT5TmemCols) : "memory");
  }
}

// ---------------------------------------------------------------------------
// FP4 form (tcgen05.mma kind::mxf4.block_scale): every descriptor bit becomes one E2M1 value, +1.0 (0x2) for a set bit and
// -1.0 (0xA) for a clear one, two per byte -- half the operand bytes of the signed-byte form and K = 64 per instruction at
// the same issue cost, which is what matters for a kernel that sits on the chip's power limit.  The block scale factors
// the instruction insists on are all 1.0 (UE8M0 0x7F): their TMEM region is filled with that byte once per CTA, so their
// layout does not matter.  Accumulators are FP32 (exact: |dot| <= 512).  Train tiles of 96 rows: four accumulators of 96
// columns leave 128 TMEM columns for the scale factors.  64-byte rows: K = 512 = two 128-byte chunks of 256 values; 48-byte rows:
// K = 384, the second chunk half zero-filled by TMA.
// ---------------------------------------------------------------------------
// Two schedules.  ALT = false: train tiles of 96 rows, two accumulator sets of two (one per query tile), the epilogue of
// tile i under the MMAs of tile i + 1 (4 x 96 columns + 128 for the scale factors).  ALT = true: train tiles of 192 rows, ONE
// accumulator per query tile; per train tile the issuer runs all K steps of query tile 0, then all of query tile 1, so that
// each accumulator is drained while the other one is computed (2 x 192 columns + 128).  The wide tile reads the 4 KB A slice
// once per 128 x 192 x 64 instruction instead of once per 128 x 96 x 64: 10 KB of shared memory per 96 tensor cycles
// instead of 7 KB per 48, which is what the narrow form is bound by.
// Instruction descriptor, block-scaled kinds (cute/arch/mma_sm100_desc.hpp InstrDescriptorBlockScaled): A = B = E2M1, which is
// format 1 for kind::mxf4 (1 << 7, 1 << 10; 5 is its code under kind::mxf8f6f4), both K-major, N >> 3 in bits 17-22, scale format UE8M0 (1 << 23), M >> 4 in bits 24-28, scale-factor ids 0, K = 64.
#ifndef BRISK_MX_RING_KB
#define BRISK_MX_RING_KB 96
#endif
constexpr int kMxRingBytes2 = BRISK_MX_RING_KB * 1024;   // B ring of the two-query-tile forms (A tiles take 64 KB)
__host__ __device__ constexpr uint32_t mx_idesc(int n) {
  return (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | (1u << 23) | ((uint32_t)(kT5M >> 4) << 24);
}

__device__ __forceinline__ void umma_mxf4(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate,
                                          uint32_t sfa_tmem, uint32_t sfb_tmem) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(sfa_tmem), "r"(sfb_tmem)
      : "memory");
}

template <int KC, int NT, bool ALT, int QT, int EW>
__global__ void __launch_bounds__((2 + EW * QT) * 32, 1)
hamming_knn2_tc5mx_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_t, long long nq,
                          long long nt, long long rows_per_split, long long train_index_offset,
                          unsigned long long* __restrict__ part, int k_bits) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int kRing = QT == 2 ? kMxRingBytes2 : kT5RingBytes;
  constexpr int kN = NT, kChunkBytes = kN * kT5Chunk, kStages = kRing / kChunkBytes;   // 24 KB (192 rows), 12 KB (96) or 16 KB (128) per stage
  constexpr int kMxColSF = (ALT ? QT : 2 * QT) * kN;   // accumulators in front, scale factors behind them
  static_assert(ALT || QT == 2, "two accumulator sets only with two query tiles");
  static_assert(EW == 4 || (EW == 8 && ALT && kN % 64 == 0), "eight epilogue warps per query tile: two per lane quadrant, half the columns each");
  static_assert(kN % 32 == 0 && kMxColSF + 128 <= 512, "the epilogue reads groups of 32 columns; the scale factors need room");
  constexpr uint32_t kIdesc = mx_idesc(kN);
  uint8_t* sA = smem;                                         // [2 tiles][KC chunks][128 rows x 128 B]
  uint8_t* sB = sA + QT * KC * kT5AChunkBytes;         // [stages][96 rows x 128 B]
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(sB + kStages * kChunkBytes);
  uint64_t* bar_empty = bar_full + kStages;
  uint64_t* bar_a = bar_empty + kStages;
  uint64_t* bar_tfull = bar_a + 1;       // [2 sets | QT tiles] accumulator ready for the epilogue
  uint64_t* bar_tempty = bar_tfull + 4;  // ... drained
  uint64_t* bar_sf = bar_tempty + 4;     // scale factors written
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_sf + 1);
  float4* s_merge = reinterpret_cast<float4*>(((uintptr_t)(tmem_slot + 1) + 15) & ~(uintptr_t)15);   // [QT * 128] (EW == 8)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long q0 = (long long)blockIdx.x * (QT * kT5M);
  const long long t_begin = (long long)blockIdx.y * rows_per_split;
  const long long t_end = min(nt, t_begin + rows_per_split);
  const int ntiles = t_end > t_begin ? (int)((t_end - t_begin + kN - 1) / kN) : 0;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
    mbar_init(bar_a, 1);
    for (int b = 0; b < (ALT ? QT : 2); ++b) { mbar_init(&bar_tfull[b], 1); mbar_init(&bar_tempty[b], ALT ? EW : 4 * QT); }
    mbar_init(bar_sf, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kT5TmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---- TMA producer ----
    if (lane == 0 && ntiles > 0) {
      mbar_expect_tx(bar_a, QT * KC * kT5AChunkBytes);
      for (int a = 0; a < QT; ++a)
        for (int c = 0; c < KC; ++c) tma_load_2d(sA + (a * KC + c) * kT5AChunkBytes, &map_q, c * kT5Chunk, (int)(q0 + a * kT5M), bar_a);
      int it = 0;
      for (int i = 0; i < ntiles; ++i)
        for (int c = 0; c < KC; ++c, ++it) {
          const int s = it % kStages;
          mbar_wait(&bar_empty[s], ((it / kStages) & 1) ^ 1);
          mbar_expect_tx(&bar_full[s], kChunkBytes);
          tma_load_2d(sB + s * kChunkBytes, &map_t, c * kT5Chunk, (int)(t_begin + (long long)i * kN), &bar_full[s]);
        }
    }
  } else if (warp == 1) {
    // ---- MMA issuer (one thread) ----
    if (lane == 0 && ntiles > 0) {
      mbar_wait(bar_a, 0);
      mbar_wait(bar_sf, 0);
      tc_fence_after();
      const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
      const uint32_t sfa = tmem_base + kMxColSF, sfb = tmem_base + kMxColSF + 64;
      // K steps of 64 values in the last chunk: 4, or 2 for 48-byte rows (the rest of that chunk is TMA's zero fill)
      const int last_steps = (k_bits - (KC - 1) * kT5Chunk * 2 + 63) / 64;
      if (ALT) {
        for (int i = 0; i < ntiles; ++i) {
#pragma unroll 1
          for (int a = 0; a < QT; ++a) {
            mbar_wait(&bar_tempty[a], (i & 1) ^ 1);
            tc_fence_after();
            for (int c = 0; c < KC; ++c) {
              const int it = i * KC + c, s = it % kStages;
              if (a == 0) { mbar_wait(&bar_full[s], (it / kStages) & 1); tc_fence_after(); }
              const int steps = c == KC - 1 ? last_steps : kT5Chunk / 32;
#pragma unroll 4
              for (int k = 0; k < steps; ++k)
                umma_mxf4(tmem_base + (uint32_t)(a * kN), umma_desc(a_base + (a * KC + c) * kT5AChunkBytes + k * 32),
                          umma_desc(b_base + s * kChunkBytes + k * 32), kIdesc, (c | k) != 0 ? 1u : 0u, sfa, sfb);
              if (a == QT - 1) umma_commit(&bar_empty[s]);   // both query tiles have read the chunk
            }
            umma_commit(&bar_tfull[a]);
          }
        }
      } else {
        int it = 0;
        for (int i = 0; i < ntiles; ++i) {
          const int b = i & 1;
          mbar_wait(&bar_tempty[b], ((i >> 1) & 1) ^ 1);
          tc_fence_after();
          for (int c = 0; c < KC; ++c, ++it) {
            const int s = it % kStages;
            mbar_wait(&bar_full[s], (it / kStages) & 1);
            tc_fence_after();
            const int steps = c == KC - 1 ? last_steps : kT5Chunk / 32;
#pragma unroll 4
            for (int k = 0; k < steps; ++k) {   // 32 bytes = 64 values per instruction
              const uint64_t bd = umma_desc(b_base + s * kChunkBytes + k * 32);
#pragma unroll
              for (int a = 0; a < QT; ++a)
                umma_mxf4(tmem_base + (uint32_t)((b * QT + a) * kN), umma_desc(a_base + (a * KC + c) * kT5AChunkBytes + k * 32), bd,
                          kIdesc, (c | k) != 0 ? 1u : 0u, sfa, sfb);
            }
            umma_commit(&bar_empty[s]);
          }
          umma_commit(&bar_tfull[b]);
        }
      }
    }
  } else {
    // ---- epilogue: thread = one query row ----
    // EW warps per query tile: one per TMEM lane quadrant (warp % 4), and with EW == 8 two per quadrant that split the columns
    const int quad = warp & 3, a = (warp - 2) / EW, half = EW == 8 ? (((warp - 2) >> 2) & 1) : 0;
    constexpr int kCols = kN / (EW / 4);   // accumulator columns a thread looks at
    const int row = a * kT5M + quad * 32 + lane;
    if (a == 0 && half == 0) {
      // scale factors: 1.0 everywhere (UE8M0 0x7f), all 128 lanes x 128 columns of the region
      uint32_t ones[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) ones[j] = 0x7f7f7f7fu;
#pragma unroll
      for (int j = 0; j < 4; ++j) tmem_st32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(kMxColSF + 32 * j), ones);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_sf);
    }
    float d0 = -1.0e9f, d1 = -1.0e9f;     // two largest dot products so far (d0 >= d1) ...
    unsigned i0 = 0xffffffffu, i1 = 0xffffffffu;  // ... and their (global) train indices
    for (int i = 0; i < ntiles; ++i) {
      const int b = ALT ? a : (i & 1);                       // barrier pair: per query tile / per accumulator set
      mbar_wait(&bar_tfull[b], ALT ? (i & 1) : ((i >> 1) & 1));
      tc_fence_after();
      const long long tile_base = t_begin + (long long)i * kN;
      const int valid = (int)min((long long)kN, t_end - tile_base);
      const unsigned idx_base = (unsigned)(train_index_offset + tile_base);
      const uint32_t acc_col = (ALT ? (uint32_t)(a * kN) : (uint32_t)((b * QT + a) * kN)) + (uint32_t)(half * kCols);
      const int col_first = half * kCols;
      // groups of 32 columns; several loads are issued before a wait so that their latencies overlap (with one load per wait
      // the epilogue of a 192-column accumulator took twice the time of the eight MMAs that hide it)
      const uint32_t acc_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc_col;
      auto process = [&](const int (&vv)[32], int cc) {
        float m = fmaxf(fmaxf(__int_as_float(vv[0]), __int_as_float(vv[1])), __int_as_float(vv[2]));
#pragma unroll
        for (int j = 3; j + 1 < 32; j += 2) m = fmaxf(fmaxf(m, __int_as_float(vv[j])), __int_as_float(vv[j + 1]));
        m = fmaxf(m, __int_as_float(vv[31]));
        if (m > d1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = col_first + cc * 32 + j;
            const float v = __int_as_float(vv[j]);
            if (v > d1 && col < valid) {
              if (v > d0) { d1 = d0; i1 = i0; d0 = v; i0 = idx_base + col; }
              else { d1 = v; i1 = idx_base + col; }
            }
          }
        }
      };
      constexpr int kGroups = kCols / 32;
      if (kGroups == 3) {
        // two loads in flight at any time (64 data registers): g0 g1 | wait | use g0, reload its registers with g2, use g1 | wait | use g2
        int va[32], vb[32];
        tmem_ld32_issue(acc_addr, va);
        tmem_ld32_issue(acc_addr + 32, vb);
        tmem_wait_ld();
        process(va, 0);
        tmem_ld32_issue(acc_addr + 64, va);
        process(vb, 1);
        tmem_wait_ld();
        process(va, 2);
      } else {
        constexpr int kBatch = kGroups % 3 == 0 ? 3 : 2;
        static_assert(kGroups % kBatch == 0, "whole batches");
#pragma unroll 1
        for (int c0 = 0; c0 < kGroups; c0 += kBatch) {
          int vi[kBatch][32];
#pragma unroll
          for (int u = 0; u < kBatch; ++u) tmem_ld32_issue(acc_addr + (uint32_t)((c0 + u) * 32), vi[u]);
          tmem_wait_ld();
#pragma unroll
          for (int u = 0; u < kBatch; ++u) process(vi[u], c0 + u);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_tempty[b]);
    }
    if (EW == 8) {
      // the two threads of a row merge their pairs: (dot, index) candidates ordered by larger dot, then lower index --
      // the order in which a single left-to-right scan would have met them
      if (half == 1) s_merge[row] = make_float4(d0, __uint_as_float(i0), d1, __uint_as_float(i1));
      asm volatile("bar.sync 1, %0;" ::"n"(EW * QT * 32) : "memory");
      if (half == 0) {
        const float4 o = s_merge[row];
        const float e0 = o.x, e1 = o.z;
        const unsigned j0 = __float_as_uint(o.y), j1 = __float_as_uint(o.w);
        auto before = [](float da, unsigned ia, float db, unsigned ib) { return da > db || (da == db && ia < ib); };
        // merge two sorted pairs (d0,i0) >= (d1,i1) and (e0,j0) >= (e1,j1)
        float r0, r1; unsigned k0, k1;
        if (before(d0, i0, e0, j0)) {
          r0 = d0; k0 = i0;
          if (before(d1, i1, e0, j0)) { r1 = d1; k1 = i1; } else { r1 = e0; k1 = j0; }
        } else {
          r0 = e0; k0 = j0;
          if (before(e1, j1, d0, i0)) { r1 = e1; k1 = j1; } else { r1 = d0; k1 = i0; }
        }
        d0 = r0; i0 = k0; d1 = r1; i1 = k1;
      }
    }
    if (half == 0 && q0 + row < nq) {
      // k_bits = descriptor bits: 512, or 384 for 48-byte rows, whose 192 expanded bytes end in the middle of the second
      // chunk -- the tensor map is 192 bytes wide and TMA fills the rest of the box with zeros, which are E2M1 0.0
      unsigned long long* out = part + ((long long)blockIdx.y * nq + q0 + row) * 2;
      out[0] = i0 == 0xffffffffu ? kT5KeyNone : ((unsigned long long)((k_bits - (int)d0) >> 1) << 32) | i0;
      out[1] = i1 == 0xffffffffu ? kT5KeyNone : ((unsigned long long)((k_bits - (int)d1) >> 1) << 32) | i1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kT5TmemCols) : "memory");
  }
}

// Descriptor bits -> E2M1 values (+1.0 = 0x2 for a set bit, -1.0 = 0xA for a clear one), bit i of a byte in nibble i:
// four output bytes per input byte.
__global__ void __launch_bounds__(256)
expand_e2m1_kernel(const uint32_t* __restrict__ src, long long n_words, uint4* __restrict__ dst) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_words) return;
  const uint32_t w = __ldg(src + i);
  uint32_t o[4];
#pragma unroll
  for (int n = 0; n < 4; ++n) o[n] = e2m1_expand_byte((w >> (8 * n)) & 0xffu);
  dst[i] = make_uint4(o[0], o[1], o[2], o[3]);
}

size_t knn_tc5mx_expanded_bytes(long long rows, int desc_bytes) {
  const long long r = rows < 256 ? 256 : rows;
  return (size_t)r * desc_bytes * 4;
}
// Query tiles of 128 rows per CTA: 2 (default), or 3 with BRISK_B200_TC5MX_QTILES=3 (384 queries share every train tile
// a CTA pulls from L2; three accumulators of 128 columns, twelve epilogue warps).
int knn_tc5mx_query_tiles() {
  static const int qt = [] { const char* e = getenv("BRISK_B200_TC5MX_QTILES"); return e && atoi(e) == 3 ? 3 : 2; }();
  return qt;
}
// Train rows per tile: 192 (default, one accumulator per query tile), 96 (BRISK_B200_TC5MX_TILE_ROWS=96: two accumulator
// sets) or 128 with three query tiles.
int knn_tc5mx_tile_rows() {
  static const int rows = [] { const char* e = getenv("BRISK_B200_TC5MX_TILE_ROWS"); return e && atoi(e) == 96 ? 96 : 192; }();
  return knn_tc5mx_query_tiles() == 3 ? 128 : rows;
}

cudaError_t launch_expand_e2m1(const uint8_t* src, long long rows, int desc_bytes, uint8_t* dst, cudaStream_t stream) {
  const long long n_words = rows * desc_bytes / 4;
  if (n_words <= 0) return cudaSuccess;
  expand_e2m1_kernel<<<(unsigned)((n_words + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const uint32_t*>(src), n_words,
                                                                              reinterpret_cast<uint4*>(dst));
  return cudaGetLastError();
}

// k == 2, 64- or 48-byte rows; map_q (128-row boxes) / map_t (tile-row boxes): tensor maps over the E2M1-expanded rows (256 / 192
// bytes each).
cudaError_t launch_hamming_knn2_tc5mx(const CUtensorMap& map_q, long long nq, const CUtensorMap& map_t, long long nt, int desc_bytes,
                                      long long train_index_offset, unsigned long long* keys, unsigned long long* part,
                                      int splits, cudaStream_t stream) {
  if (nq <= 0) return cudaSuccess;
  if (desc_bytes != 64 && desc_bytes != 48) return cudaErrorInvalidValue;
  if (nt <= 0) return cudaMemsetAsync(keys, 0xff, (size_t)nq * 2 * 8, stream);
  const int tile = knn_tc5mx_tile_rows(), qt = knn_tc5mx_query_tiles();
  // epilogue warps per query tile: 4, one per TMEM lane quadrant.  BRISK_B200_TC5MX_EPI_WARPS=8 selects two per quadrant (half the
  // columns each, pairs merged through shared memory): measured 4.2 against 4.6 Tcmp/s -- the epilogue is bound by what TMEM
  // delivers (every 4-byte accumulator is read once: 64 B per cycle and SM is 4.6 T comparisons per second), not by its warps.
  static const int epi_env = [] { const char* e = getenv("BRISK_B200_TC5MX_EPI_WARPS"); return e ? atoi(e) : 0; }();
  const int ew = (qt == 2 && tile == 192 && epi_env == 8) ? 8 : 4;
  long long rows_per_split = ((nt + splits - 1) / splits + tile - 1) / tile * tile;
  if (rows_per_split <= 0) rows_per_split = tile;
  unsigned long long* dst = splits == 1 ? keys : part;
  constexpr int KC = 2;
  const size_t smem = (size_t)qt * KC * kT5AChunkBytes + (qt == 2 ? kMxRingBytes2 : kT5RingBytes) + 1024 /* alignment */ + 512 /* barriers */ +
                      4096 + 64 /* merge scratch of the eight-warp epilogue */;
  dim3 grid((unsigned)((nq + qt * kT5M - 1) / (qt * kT5M)), splits);
  const int threads = (2 + ew * qt) * 32;
  cudaError_t e;
  if (qt == 3) {
    e = cudaFuncSetAttribute(hamming_knn2_tc5mx_kernel<KC, 128, true, 3, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    hamming_knn2_tc5mx_kernel<KC, 128, true, 3, 4><<<grid, threads, smem, stream>>>(map_q, map_t, nq, nt, rows_per_split, train_index_offset, dst, desc_bytes * 8);
  } else if (tile == 192 && ew == 8) {
    e = cudaFuncSetAttribute(hamming_knn2_tc5mx_kernel<KC, 192, true, 2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    hamming_knn2_tc5mx_kernel<KC, 192, true, 2, 8><<<grid, threads, smem, stream>>>(map_q, map_t, nq, nt, rows_per_split, train_index_offset, dst, desc_bytes * 8);
  } else if (tile == 192) {
    e = cudaFuncSetAttribute(hamming_knn2_tc5mx_kernel<KC, 192, true, 2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    hamming_knn2_tc5mx_kernel<KC, 192, true, 2, 4><<<grid, threads, smem, stream>>>(map_q, map_t, nq, nt, rows_per_split, train_index_offset, dst, desc_bytes * 8);
  } else {
    e = cudaFuncSetAttribute(hamming_knn2_tc5mx_kernel<KC, 96, false, 2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    hamming_knn2_tc5mx_kernel<KC, 96, false, 2, 4><<<grid, threads, smem, stream>>>(map_q, map_t, nq, nt, rows_per_split, train_index_offset, dst, desc_bytes * 8);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (splits > 1) e = launch_knn_merge(part, splits, nq, 2, keys, stream);
  return e;
}

// Descriptor bits -> signed bytes (+1 / -1), 32 per input word.
__global__ void __launch_bounds__(256)
expand_pm1_kernel(const uint32_t* __restrict__ src, long long n_words, uint4* __restrict__ dst) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_words) return;
  const uint32_t w = __ldg(src + i);
  uint32_t o[8];
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    // bit j of the nibble -> top bit of byte j (multiply), replicated over the byte (PRMT sign mode): 0x00 / 0xff
    uint32_t mask;
    asm("prmt.b32 %0, %1, %1, 0xba98;" : "=r"(mask) : "r"(((w >> (4 * n)) & 0xfu) * 0x10204080u));
    o[n] = (mask & 0x01010101u) | ~mask;   // set -> 0x01, clear -> 0xff
  }
  dst[2 * i] = make_uint4(o[0], o[1], o[2], o[3]);
  dst[2 * i + 1] = make_uint4(o[4], o[5], o[6], o[7]);
}

size_t knn_tc5_expanded_bytes(long long rows, int desc_bytes) {
  const long long r = rows < 256 ? 256 : rows;   // the tensor map's row extent is at least one box
  return (size_t)r * desc_bytes * 8;
}

cudaError_t launch_expand_pm1(const uint8_t* src, long long rows, int desc_bytes, uint8_t* dst, cudaStream_t stream) {
  const long long n_words = rows * desc_bytes / 4;
  if (n_words <= 0) return cudaSuccess;
  expand_pm1_kernel<<<(unsigned)((n_words + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const uint32_t*>(src), n_words,
                                                                             reinterpret_cast<uint4*>(dst));
  return cudaGetLastError();
}

int knn_tc5_tile_rows() { return t5_tile_rows(); }

// Splits of the train set: enough CTAs for the 148 SMs, then the smallest count whose last wave is >= 95 % full.
int knn_tc5_num_splits(long long nq, long long nt, int query_tiles) {
  if (query_tiles <= 0) query_tiles = kT5QTiles;
  const long long qblocks = (nq + query_tiles * kT5M - 1) / (query_tiles * kT5M);
  long long max_splits = (nt + 16 * 256 - 1) / (16 * 256);
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

template <int KC, int NT>
static cudaError_t launch_tc5(const CUtensorMap& mq, const CUtensorMap& mt, long long nq, long long nt, long long off,
                              unsigned long long* dst, int splits, long long rows_per_split, cudaStream_t stream) {
  const size_t smem = (size_t)kT5QTiles * KC * kT5AChunkBytes + kT5RingBytes + 1024 /* alignment */ + 256 /* barriers */;
  cudaError_t e = cudaFuncSetAttribute(hamming_knn2_tc5_kernel<KC, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  dim3 grid((unsigned)((nq + kT5QTiles * kT5M - 1) / (kT5QTiles * kT5M)), splits);
  hamming_knn2_tc5_kernel<KC, NT><<<grid, kT5Threads, smem, stream>>>(mq, mt, nq, nt, rows_per_split, off, dst);
  return cudaGetLastError();
}

// 0 (default): queries in shared memory (hamming_knn2_tc5_kernel); 1: queries in TMEM (hamming_knn2_tc5ts_kernel,
// BRISK_B200_TC5_MODE=ts in the environment).  Measured on B200 (100 k x 1 M, k = 2): 2.76 / 2.77 Tcmp/s at 512 bit -- both
// run into the chip's power limit (SM clock ~1.33 GHz under sustained tensor load, MEASURED_PEAKS.json) -- and 4.18 / 3.37
// Tcmp/s at 384 bit, where the TMEM form's single accumulator per tile leaves the epilogue too little time.
int knn_tc5_queries_in_tmem() {
  static const int ts = [] { const char* e = getenv("BRISK_B200_TC5_MODE"); return e && e[0] == 't' && e[1] == 's' ? 1 : 0; }();
  return ts;
}

template <int KC>
static cudaError_t launch_tc5ts(const uint8_t* q, const CUtensorMap& mt, long long nq, long long nt, long long off,
                                unsigned long long* dst, int splits, long long rows_per_split, cudaStream_t stream) {
  const size_t smem = (size_t)kTsStages * 128 * kT5Chunk + 1024 /* alignment */ + 512 /* barriers */;
  cudaError_t e = cudaFuncSetAttribute(hamming_knn2_tc5ts_kernel<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  dim3 grid((unsigned)((nq + kT5QTiles * kT5M - 1) / (kT5QTiles * kT5M)), splits);
  hamming_knn2_tc5ts_kernel<KC><<<grid, kT5Threads, smem, stream>>>(q, mt, nq, nt, rows_per_split, off, dst);
  return cudaGetLastError();
}

// Queries as raw descriptor rows (device, 16-byte aligned), train rows expanded + tensor map with a 128-row box.
cudaError_t launch_hamming_knn2_tc5ts(const uint8_t* q, long long nq, const CUtensorMap& map_t, long long nt, int desc_bytes,
                                      long long train_index_offset, unsigned long long* keys, unsigned long long* part,
                                      int splits, cudaStream_t stream) {
  if (nq <= 0) return cudaSuccess;
  if (nt <= 0) return cudaMemsetAsync(keys, 0xff, (size_t)nq * 2 * 8, stream);
  long long rows_per_split = ((nt + splits - 1) / splits + 127) / 128 * 128;
  if (rows_per_split <= 0) rows_per_split = 128;
  unsigned long long* dst = splits == 1 ? keys : part;
  cudaError_t e;
  if (desc_bytes == 64) e = launch_tc5ts<4>(q, map_t, nq, nt, train_index_offset, dst, splits, rows_per_split, stream);
  else if (desc_bytes == 48) e = launch_tc5ts<3>(q, map_t, nq, nt, train_index_offset, dst, splits, rows_per_split, stream);
  else return cudaErrorInvalidValue;
  if (e != cudaSuccess) return e;
  if (splits > 1) e = launch_knn_merge(part, splits, nq, 2, keys, stream);
  return e;
}

// k == 2 only; map_q / map_t: tensor maps over the EXPANDED rows (knn_tc5_encode_map in capi.cu); keys [nq][2];
// part: scratch [splits][nq][2] (unused when splits == 1).
cudaError_t launch_hamming_knn2_tc5(const CUtensorMap& map_q, long long nq, const CUtensorMap& map_t, long long nt, int desc_bytes,
                                    long long train_index_offset, unsigned long long* keys, unsigned long long* part,
                                    int splits, cudaStream_t stream) {
  if (nq <= 0) return cudaSuccess;
  const int tile = t5_tile_rows();
  long long rows_per_split = ((nt + splits - 1) / splits + tile - 1) / tile * tile;
  if (rows_per_split <= 0) rows_per_split = tile;
  unsigned long long* dst = splits == 1 ? keys : part;
  cudaError_t e;
  if (nt <= 0) {
    e = cudaMemsetAsync(keys, 0xff, (size_t)nq * 2 * 8, stream);   // no train rows: every key is "none"
    return e;
  }
  if (desc_bytes == 64 && tile == 256) e = launch_tc5<4, 256>(map_q, map_t, nq, nt, train_index_offset, dst, splits, rows_per_split, stream);
  else if (desc_bytes == 48 && tile == 256) e = launch_tc5<3, 256>(map_q, map_t, nq, nt, train_index_offset, dst, splits, rows_per_split, stream);
  else if (desc_bytes == 64) e = launch_tc5<4, 128>(map_q, map_t, nq, nt, train_index_offset, dst, splits, rows_per_split, stream);
  else if (desc_bytes == 48) e = launch_tc5<3, 128>(map_q, map_t, nq, nt, train_index_offset, dst, splits, rows_per_split, stream);
  else return cudaErrorInvalidValue;
  if (e != cudaSuccess) return e;
  if (splits > 1) e = launch_knn_merge(part, splits, nq, 2, keys, stream);
  return e;
}

}  // namespace briskb200
