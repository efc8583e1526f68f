// Brute-force Hamming k-nearest-neighbour search (sm_100a).
//
// Replaces brisk::Hamming (reference brisk/include/brisk/internal/hamming-inl.h:
// 85-134, SSSE3 nibble-LUT popcount of the XOR) and the k successive arg-min
// passes of BruteForceMatcher::commonKnnMatchImpl (brisk/src/brute-force-matcher.cc:
// 80-162).  Each thread keeps two query descriptors in registers; train rows are
// staged in shared memory and read as warp-wide broadcasts, so a row costs one
// LDS.128 per 16 bytes and the kernel is bound by the POPC pipe.  Candidates are
// (distance << 32 | train index) keys: the unsigned minimum of a key is the
// nearest neighbour with the lowest train index on ties, which is the
// reference's tie rule (first minimum of a left-to-right scan wins).
#include <cuda_runtime.h>

#include "gcc_sort.cuh"
#include "kernels.h"

namespace briskb200 {

constexpr int kKnnThreads = 256;
constexpr int kKnnQPerThread = 2;
constexpr int kKnnTile = 128;  // train rows per shared-memory tile
constexpr unsigned long long kKeyNone = ~0ull;

template <int K>
__device__ __forceinline__ void topk_insert(unsigned long long (&best)[K], unsigned long long key) {
  if (key < best[K - 1]) {
    best[K - 1] = key;
#pragma unroll
    for (int i = K - 1; i > 0; --i) {
      if (best[i] < best[i - 1]) { const unsigned long long t = best[i]; best[i] = best[i - 1]; best[i - 1] = t; }
    }
  }
}

// MASKED: `mask` is the [nq][nt] byte matrix of DescriptorMatcher's masks (0 = pair not allowed,
// brute-force-matcher.cc:118-119); excluded pairs never become candidates.
template <int WORDS, int K, bool MASKED = false>
__global__ void __launch_bounds__(kKnnThreads)
hamming_knn_kernel(const uint32_t* __restrict__ q, long long nq, const uint32_t* __restrict__ t, long long nt,
                   long long rows_per_split, long long train_index_offset, unsigned long long* __restrict__ part,
                   const uint8_t* __restrict__ mask = nullptr) {
  __shared__ __align__(16) uint32_t s_t[kKnnTile * WORDS];
  const int tid = threadIdx.x;
  const long long q0 = ((long long)blockIdx.x * kKnnThreads + tid) * kKnnQPerThread;
  uint32_t qa[WORDS], qb[WORDS];
#pragma unroll
  for (int i = 0; i < WORDS; ++i) {
    qa[i] = q0 < nq ? q[q0 * WORDS + i] : 0u;
    qb[i] = q0 + 1 < nq ? q[(q0 + 1) * WORDS + i] : 0u;
  }
  unsigned long long ba[K], bb[K];
#pragma unroll
  for (int i = 0; i < K; ++i) { ba[i] = kKeyNone; bb[i] = kKeyNone; }

  const long long t_begin = (long long)blockIdx.y * rows_per_split;
  const long long t_end = min(nt, t_begin + rows_per_split);
  for (long long base = t_begin; base < t_end; base += kKnnTile) {
    const int rows = (int)min((long long)kKnnTile, t_end - base);
    __syncthreads();
    // stage the tile: 16-byte chunks, coalesced
    const uint4* src = reinterpret_cast<const uint4*>(t + base * WORDS);
    uint4* dst = reinterpret_cast<uint4*>(s_t);
    for (int i = tid; i < rows * (WORDS / 4); i += kKnnThreads) dst[i] = src[i];
    __syncthreads();
    for (int r = 0; r < rows; ++r) {
      const uint4* row = reinterpret_cast<const uint4*>(s_t + r * WORDS);
      int da = 0, db = 0;
#pragma unroll
      for (int c = 0; c < WORDS / 4; ++c) {
        const uint4 v = row[c];  // broadcast
        da += __popc(qa[4 * c] ^ v.x) + __popc(qa[4 * c + 1] ^ v.y) + __popc(qa[4 * c + 2] ^ v.z) + __popc(qa[4 * c + 3] ^ v.w);
        db += __popc(qb[4 * c] ^ v.x) + __popc(qb[4 * c + 1] ^ v.y) + __popc(qb[4 * c + 2] ^ v.z) + __popc(qb[4 * c + 3] ^ v.w);
      }
      const unsigned long long idx = (unsigned long long)(train_index_offset + base + r);
      if (!MASKED || (q0 < nq && mask[q0 * nt + base + r])) topk_insert<K>(ba, ((unsigned long long)(uint32_t)da << 32) | idx);
      if (!MASKED || (q0 + 1 < nq && mask[(q0 + 1) * nt + base + r])) topk_insert<K>(bb, ((unsigned long long)(uint32_t)db << 32) | idx);
    }
  }
  unsigned long long* out = part + (long long)blockIdx.y * nq * K;
  if (q0 < nq) {
#pragma unroll
    for (int i = 0; i < K; ++i) out[q0 * K + i] = ba[i];
  }
  if (q0 + 1 < nq) {
#pragma unroll
    for (int i = 0; i < K; ++i) out[(q0 + 1) * K + i] = bb[i];
  }
}

// k-way merge of `n_lists` sorted key lists per query ([list][query][k]) into
// out[query][k]; also the NCCL top-k merge of per-GPU shards.
__global__ void __launch_bounds__(256)
knn_merge_kernel(const unsigned long long* __restrict__ lists, int n_lists, long long nq, int k, unsigned long long* __restrict__ out) {
  const long long qi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= nq) return;
  unsigned long long best[8];
  for (int i = 0; i < 8; ++i) best[i] = kKeyNone;
  for (int l = 0; l < n_lists; ++l) {
    const unsigned long long* src = lists + ((long long)l * nq + qi) * k;
    for (int i = 0; i < k; ++i) {
      const unsigned long long key = src[i];
      if (key >= best[k - 1]) break;  // lists are sorted
      best[k - 1] = key;
      for (int j = k - 1; j > 0 && best[j] < best[j - 1]; --j) { const unsigned long long tmp = best[j]; best[j] = best[j - 1]; best[j - 1] = tmp; }
    }
  }
  for (int i = 0; i < k; ++i) out[qi * k + i] = best[i];
}

__global__ void __launch_bounds__(256)
knn_unpack_kernel(const unsigned long long* __restrict__ keys, long long n, int32_t* __restrict__ idx, int32_t* __restrict__ dist) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long key = keys[i];
  if (key == kKeyNone) { idx[i] = -1; dist[i] = -1; }
  else { idx[i] = (int32_t)(key & 0xffffffffull); dist[i] = (int32_t)(key >> 32); }
}

template <int WORDS>
static cudaError_t launch_knn_words(const uint8_t* q, long long nq, const uint8_t* t, long long nt, int k, long long off,
                                    unsigned long long* keys, unsigned long long* part, int splits, long long rows_per_split,
                                    cudaStream_t stream) {
  dim3 grid((unsigned)((nq + kKnnThreads * kKnnQPerThread - 1) / (kKnnThreads * kKnnQPerThread)), splits);
  unsigned long long* dst = splits == 1 ? keys : part;
  const uint32_t* q32 = reinterpret_cast<const uint32_t*>(q);
  const uint32_t* t32 = reinterpret_cast<const uint32_t*>(t);
  switch (k) {
    case 1: hamming_knn_kernel<WORDS, 1><<<grid, kKnnThreads, 0, stream>>>(q32, nq, t32, nt, rows_per_split, off, dst); break;
    case 2: hamming_knn_kernel<WORDS, 2><<<grid, kKnnThreads, 0, stream>>>(q32, nq, t32, nt, rows_per_split, off, dst); break;
    case 3: case 4: hamming_knn_kernel<WORDS, 4><<<grid, kKnnThreads, 0, stream>>>(q32, nq, t32, nt, rows_per_split, off, dst); break;
    default: hamming_knn_kernel<WORDS, 8><<<grid, kKnnThreads, 0, stream>>>(q32, nq, t32, nt, rows_per_split, off, dst); break;
  }
  return cudaGetLastError();
}

template <int WORDS>
static cudaError_t launch_knn_masked_words(const uint8_t* q, long long nq, const uint8_t* t, long long nt, int k, const uint8_t* mask,
                                           unsigned long long* keys, cudaStream_t stream) {
  dim3 grid((unsigned)((nq + kKnnThreads * kKnnQPerThread - 1) / (kKnnThreads * kKnnQPerThread)), 1);
  const uint32_t* q32 = reinterpret_cast<const uint32_t*>(q);
  const uint32_t* t32 = reinterpret_cast<const uint32_t*>(t);
  const long long rows = (nt + kKnnTile - 1) / kKnnTile * kKnnTile;
  switch (k) {
    case 1: hamming_knn_kernel<WORDS, 1, true><<<grid, kKnnThreads, 0, stream>>>(q32, nq, t32, nt, rows, 0, keys, mask); break;
    case 2: hamming_knn_kernel<WORDS, 2, true><<<grid, kKnnThreads, 0, stream>>>(q32, nq, t32, nt, rows, 0, keys, mask); break;
    case 3: case 4: hamming_knn_kernel<WORDS, 4, true><<<grid, kKnnThreads, 0, stream>>>(q32, nq, t32, nt, rows, 0, keys, mask); break;
    default: hamming_knn_kernel<WORDS, 8, true><<<grid, kKnnThreads, 0, stream>>>(q32, nq, t32, nt, rows, 0, keys, mask); break;
  }
  return cudaGetLastError();
}

// kNN with a [nq][nt] byte mask; keys [nq][knn_round_k(k)].
cudaError_t launch_hamming_knn_masked(const uint8_t* q, long long nq, const uint8_t* t, long long nt, int desc_bytes, int k,
                                      const uint8_t* mask, unsigned long long* keys, cudaStream_t stream) {
  if (nq <= 0) return cudaSuccess;
  const int kr = knn_round_k(k);
  switch (desc_bytes) {
    case 48: return launch_knn_masked_words<12>(q, nq, t, nt, kr, mask, keys, stream);
    case 64: return launch_knn_masked_words<16>(q, nq, t, nt, kr, mask, keys, stream);
    case 128: return launch_knn_masked_words<32>(q, nq, t, nt, kr, mask, keys, stream);
    default: return cudaErrorInvalidValue;
  }
}

// ---------------------------------------------------------------------------
// radiusMatch: BruteForceMatcher::commonRadiusMatchImpl (reference
// brisk/src/brute-force-matcher.cc:164-214).  Two sweeps with the same tiling
// as the kNN kernel, one thread per query: count the train rows with
// (float)d < maxDistance (and an allowing mask byte), prefix-sum the counts, then
// emit (train index, distance) pairs in train order -- the order in which the
// reference pushes them -- and finally std::sort each query's list by distance the
// way libstdc++ does (gcc_sort.cuh: equal distances keep the reference's order).
// ---------------------------------------------------------------------------
struct RadiusMatch { int idx, dist; };
struct RadiusLess { BRISK_HD bool operator()(const RadiusMatch& a, const RadiusMatch& b) const { return a.dist < b.dist; } };

template <int WORDS, bool EMIT>
__global__ void __launch_bounds__(kKnnThreads)
hamming_radius_kernel(const uint32_t* __restrict__ q, long long nq, const uint32_t* __restrict__ t, long long nt, float max_distance,
                      const uint8_t* __restrict__ mask, long long* __restrict__ counts, const long long* __restrict__ offsets,
                      RadiusMatch* __restrict__ out, long long capacity) {
  __shared__ __align__(16) uint32_t s_t[kKnnTile * WORDS];
  const int tid = threadIdx.x;
  const long long qi = (long long)blockIdx.x * kKnnThreads + tid;
  uint32_t qa[WORDS];
#pragma unroll
  for (int i = 0; i < WORDS; ++i) qa[i] = qi < nq ? q[qi * WORDS + i] : 0u;
  long long n = 0;
  const long long o = EMIT && qi < nq ? offsets[qi] : 0;
  for (long long base = 0; base < nt; base += kKnnTile) {
    const int rows = (int)min((long long)kKnnTile, nt - base);
    __syncthreads();
    const uint4* src = reinterpret_cast<const uint4*>(t + base * WORDS);
    uint4* dst = reinterpret_cast<uint4*>(s_t);
    for (int i = tid; i < rows * (WORDS / 4); i += kKnnThreads) dst[i] = src[i];
    __syncthreads();
    if (qi >= nq) continue;
    for (int r = 0; r < rows; ++r) {
      const uint4* row = reinterpret_cast<const uint4*>(s_t + r * WORDS);
      int d = 0;
#pragma unroll
      for (int c = 0; c < WORDS / 4; ++c) {
        const uint4 v = row[c];
        d += __popc(qa[4 * c] ^ v.x) + __popc(qa[4 * c + 1] ^ v.y) + __popc(qa[4 * c + 2] ^ v.z) + __popc(qa[4 * c + 3] ^ v.w);
      }
      if ((float)d < max_distance && (!mask || mask[qi * nt + base + r])) {
        if (EMIT && o + n < capacity) out[o + n] = RadiusMatch{(int)(base + r), d};
        ++n;
      }
    }
  }
  if (!EMIT && qi < nq) counts[qi] = n;
}

// exclusive prefix sum of counts[n] into offsets[n + 1] (one CTA; n is a query count)
__global__ void __launch_bounds__(1024)
radius_scan_kernel(const long long* __restrict__ counts, long long n, long long* __restrict__ offsets) {
  __shared__ long long s_warp[32];
  __shared__ long long s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (long long base = 0; base < n; base += 1024) {
    const long long i = base + tid;
    const long long v = i < n ? counts[i] : 0;
    long long x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const long long y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    long long before = s_carry;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    if (i < n) offsets[i] = before + x - v;
    __syncthreads();
    if (tid == 1023) s_carry = before + x;
    __syncthreads();
  }
  if (tid == 0) offsets[n] = s_carry;
}

__global__ void __launch_bounds__(128)
radius_sort_kernel(RadiusMatch* __restrict__ m, const long long* __restrict__ offsets, long long nq, long long capacity) {
  const long long qi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= nq) return;
  const long long b = offsets[qi], e = offsets[qi + 1];
  if (e > capacity || e - b < 2) return;
  gs_sort(RadiusLess(), m + b, (int)(e - b));
}

__global__ void __launch_bounds__(256)
radius_unpack_kernel(const RadiusMatch* __restrict__ m, long long n, int32_t* __restrict__ idx, int32_t* __restrict__ dist) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  idx[i] = m[i].idx; dist[i] = m[i].dist;
}

// ---------------------------------------------------------------------------
// Any row width, any k.  The reference's matcher takes whatever cv::Mat it is given: rows of `cols` bytes, of which
// brisk::Hamming compares cols / 16 whole 128-bit words (hamming.h:101-113), and any k (brute-force-matcher.cc:80-162).
// The kernels above are instantiated for the extractors' widths (48 / 64 / 128 bytes) and k <= 8; everything else goes
// through these: the queries of a CTA and a tile of train rows in shared memory, one thread per query, the k best in
// passes of eight (pass p keeps the eight smallest keys above the last key of pass p - 1; keys are unique, they
// carry the train index).  Rows must be a multiple of 4 bytes, at most 496.
// ---------------------------------------------------------------------------
constexpr int kAnyThreads = 128, kAnyTile = 32;

__global__ void __launch_bounds__(kAnyThreads)
hamming_knn_any_kernel(const uint32_t* __restrict__ q, long long nq, const uint32_t* __restrict__ t, long long nt, int row_words,
                       int cmp_words, long long train_index_offset, unsigned long long* __restrict__ keys, int kr, int pass,
                       const uint8_t* __restrict__ mask) {
  extern __shared__ __align__(16) uint32_t s_any[];
  uint32_t* s_q = s_any;                                        // [kAnyThreads][cmp_words + 1]
  uint32_t* s_t = s_any + kAnyThreads * (cmp_words + 1);        // [kAnyTile][cmp_words]
  const int tid = threadIdx.x;
  const long long q_first = (long long)blockIdx.x * kAnyThreads, qi = q_first + tid;
  for (int i = tid; i < kAnyThreads * cmp_words; i += kAnyThreads) {
    const int r = i / cmp_words, c = i - r * cmp_words;
    s_q[r * (cmp_words + 1) + c] = q_first + r < nq ? q[(q_first + r) * row_words + c] : 0u;
  }
  unsigned long long best[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) best[i] = kKeyNone;
  const bool bounded = pass > 0;
  const unsigned long long lower = (bounded && qi < nq) ? keys[qi * kr + 8 * pass - 1] : 0ull;
  const uint32_t* mine = s_q + tid * (cmp_words + 1);
  for (long long base = 0; base < nt; base += kAnyTile) {
    const int rows = (int)min((long long)kAnyTile, nt - base);
    __syncthreads();
    for (int i = tid; i < rows * cmp_words; i += kAnyThreads) {
      const int r = i / cmp_words, c = i - r * cmp_words;
      s_t[i] = t[(base + r) * row_words + c];
    }
    __syncthreads();
    if (qi >= nq || (bounded && lower == kKeyNone)) continue;   // (no more candidates after a pass that ran short)
    for (int r = 0; r < rows; ++r) {
      int d = 0;
      for (int c = 0; c < cmp_words; ++c) d += __popc(mine[c] ^ s_t[r * cmp_words + c]);
      const unsigned long long key = ((unsigned long long)(uint32_t)d << 32) | (unsigned long long)(train_index_offset + base + r);
      if (mask && !mask[qi * nt + base + r]) continue;
      if (bounded && key <= lower) continue;
      topk_insert<8>(best, key);
    }
  }
  if (qi < nq) {
#pragma unroll
    for (int i = 0; i < 8; ++i) keys[qi * kr + 8 * pass + i] = best[i];
  }
}

int knn_any_round_k(int k) { return (k + 7) / 8 * 8; }

// keys [nq][knn_any_round_k(k)], ascending per query, kKeyNone where a query runs out of (allowed) train rows.
cudaError_t launch_hamming_knn_any(const uint8_t* q, long long nq, const uint8_t* t, long long nt, int desc_bytes, int k,
                                   long long train_index_offset, const uint8_t* mask, unsigned long long* keys, cudaStream_t stream) {
  if (nq <= 0) return cudaSuccess;
  if (desc_bytes < 4 || desc_bytes % 4 || desc_bytes > 496 || k < 1) return cudaErrorInvalidValue;
  const int row_words = desc_bytes / 4, cmp_words = (desc_bytes / 16) * 4, kr = knn_any_round_k(k);
  const size_t smem = ((size_t)kAnyThreads * (cmp_words + 1) + (size_t)kAnyTile * cmp_words) * 4;
  cudaError_t e = cudaFuncSetAttribute(hamming_knn_any_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const unsigned grid = (unsigned)((nq + kAnyThreads - 1) / kAnyThreads);
  for (int pass = 0; pass < kr / 8; ++pass)
    hamming_knn_any_kernel<<<grid, kAnyThreads, smem, stream>>>(reinterpret_cast<const uint32_t*>(q), nq, reinterpret_cast<const uint32_t*>(t), nt,
                                                                row_words, cmp_words, train_index_offset, keys, kr, pass, mask);
  return cudaGetLastError();
}

template <bool EMIT>
__global__ void __launch_bounds__(kAnyThreads)
hamming_radius_any_kernel(const uint32_t* __restrict__ q, long long nq, const uint32_t* __restrict__ t, long long nt, int row_words,
                          int cmp_words, float max_distance, const uint8_t* __restrict__ mask, long long* __restrict__ counts,
                          const long long* __restrict__ offsets, RadiusMatch* __restrict__ out, long long capacity) {
  extern __shared__ __align__(16) uint32_t s_any[];
  uint32_t* s_q = s_any;
  uint32_t* s_t = s_any + kAnyThreads * (cmp_words + 1);
  const int tid = threadIdx.x;
  const long long q_first = (long long)blockIdx.x * kAnyThreads, qi = q_first + tid;
  for (int i = tid; i < kAnyThreads * cmp_words; i += kAnyThreads) {
    const int r = i / cmp_words, c = i - r * cmp_words;
    s_q[r * (cmp_words + 1) + c] = q_first + r < nq ? q[(q_first + r) * row_words + c] : 0u;
  }
  const uint32_t* mine = s_q + tid * (cmp_words + 1);
  long long n = 0;
  const long long o = EMIT && qi < nq ? offsets[qi] : 0;
  for (long long base = 0; base < nt; base += kAnyTile) {
    const int rows = (int)min((long long)kAnyTile, nt - base);
    __syncthreads();
    for (int i = tid; i < rows * cmp_words; i += kAnyThreads) {
      const int r = i / cmp_words, c = i - r * cmp_words;
      s_t[i] = t[(base + r) * row_words + c];
    }
    __syncthreads();
    if (qi >= nq) continue;
    for (int r = 0; r < rows; ++r) {
      int d = 0;
      for (int c = 0; c < cmp_words; ++c) d += __popc(mine[c] ^ s_t[r * cmp_words + c]);
      if ((float)d < max_distance && (!mask || mask[qi * nt + base + r])) {
        if (EMIT && o + n < capacity) out[o + n] = RadiusMatch{(int)(base + r), d};
        ++n;
      }
    }
  }
  if (!EMIT && qi < nq) counts[qi] = n;
}

static cudaError_t radius_any(int phase, const uint8_t* q, long long nq, const uint8_t* t, long long nt, int desc_bytes, float max_distance,
                              const uint8_t* mask, long long* counts, long long* offsets, RadiusMatch* out, long long capacity,
                              cudaStream_t stream) {
  if (desc_bytes < 4 || desc_bytes % 4 || desc_bytes > 496) return cudaErrorInvalidValue;
  const int row_words = desc_bytes / 4, cmp_words = (desc_bytes / 16) * 4;
  const size_t smem = ((size_t)kAnyThreads * (cmp_words + 1) + (size_t)kAnyTile * cmp_words) * 4;
  const unsigned grid = (unsigned)((nq + kAnyThreads - 1) / kAnyThreads);
  const uint32_t* q32 = reinterpret_cast<const uint32_t*>(q);
  const uint32_t* t32 = reinterpret_cast<const uint32_t*>(t);
  cudaError_t e = cudaFuncSetAttribute(hamming_radius_any_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(hamming_radius_any_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (phase == 0) hamming_radius_any_kernel<false><<<grid, kAnyThreads, smem, stream>>>(q32, nq, t32, nt, row_words, cmp_words, max_distance, mask, counts, offsets, out, capacity);
  else hamming_radius_any_kernel<true><<<grid, kAnyThreads, smem, stream>>>(q32, nq, t32, nt, row_words, cmp_words, max_distance, mask, counts, offsets, out, capacity);
  return cudaGetLastError();
}

template <int WORDS>
static cudaError_t radius_words(int phase, const uint8_t* q, long long nq, const uint8_t* t, long long nt, float max_distance,
                                const uint8_t* mask, long long* counts, long long* offsets, RadiusMatch* out, long long capacity,
                                cudaStream_t stream) {
  const unsigned grid = (unsigned)((nq + kKnnThreads - 1) / kKnnThreads);
  const uint32_t* q32 = reinterpret_cast<const uint32_t*>(q);
  const uint32_t* t32 = reinterpret_cast<const uint32_t*>(t);
  if (phase == 0) hamming_radius_kernel<WORDS, false><<<grid, kKnnThreads, 0, stream>>>(q32, nq, t32, nt, max_distance, mask, counts, offsets, out, capacity);
  else hamming_radius_kernel<WORDS, true><<<grid, kKnnThreads, 0, stream>>>(q32, nq, t32, nt, max_distance, mask, counts, offsets, out, capacity);
  return cudaGetLastError();
}

// phase 0: counts[nq] and offsets[nq + 1]; phase 1: matches (8 bytes each: train index, distance) at
// offsets, up to `capacity` in all, optionally sorted per query the way std::sort does.
cudaError_t launch_hamming_radius(int phase, const uint8_t* q, long long nq, const uint8_t* t, long long nt, int desc_bytes,
                                  float max_distance, const uint8_t* mask, long long* counts, long long* offsets, void* matches,
                                  long long capacity, int sort, cudaStream_t stream) {
  if (nq <= 0) return cudaSuccess;
  RadiusMatch* out = static_cast<RadiusMatch*>(matches);
  cudaError_t e;
  switch (desc_bytes) {
    case 48: e = radius_words<12>(phase, q, nq, t, nt, max_distance, mask, counts, offsets, out, capacity, stream); break;
    case 64: e = radius_words<16>(phase, q, nq, t, nt, max_distance, mask, counts, offsets, out, capacity, stream); break;
    case 128: e = radius_words<32>(phase, q, nq, t, nt, max_distance, mask, counts, offsets, out, capacity, stream); break;
    default: e = radius_any(phase, q, nq, t, nt, desc_bytes, max_distance, mask, counts, offsets, out, capacity, stream); break;
  }
  if (e != cudaSuccess) return e;
  if (phase == 0) radius_scan_kernel<<<1, 1024, 0, stream>>>(counts, nq, offsets);
  else if (sort) radius_sort_kernel<<<(unsigned)((nq + 127) / 128), 128, 0, stream>>>(out, offsets, nq, capacity);
  return cudaGetLastError();
}

cudaError_t launch_radius_unpack(const void* matches, long long n, int32_t* idx, int32_t* dist, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  radius_unpack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(static_cast<const RadiusMatch*>(matches), n, idx, dist);
  return cudaGetLastError();
}

// Number of train-set splits used so that small query sets still fill the GPU.
int knn_num_splits(long long nq, long long nt) {
  const long long qblocks = (nq + kKnnThreads * kKnnQPerThread - 1) / (kKnnThreads * kKnnQPerThread);
  long long splits = (2 * 148 + qblocks - 1) / qblocks;
  const long long max_splits = (nt + 4 * kKnnTile - 1) / (4 * kKnnTile);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits > 1024) splits = 1024;
  return (int)splits;
}

int knn_round_k(int k) { return k <= 1 ? 1 : (k <= 2 ? 2 : (k <= 4 ? 4 : 8)); }

// keys: [nq][kr] (kr = knn_round_k(k)); part: scratch [splits][nq][kr] (unused when splits == 1).
cudaError_t launch_hamming_knn_ex(const uint8_t* q, long long nq, const uint8_t* t, long long nt, int desc_bytes, int k,
                                  long long train_index_offset, unsigned long long* keys, unsigned long long* part,
                                  int splits, cudaStream_t stream) {
  if (nq <= 0) return cudaSuccess;
  const int kr = knn_round_k(k);
  const long long rows_per_split = ((nt + splits - 1) / splits + kKnnTile - 1) / kKnnTile * kKnnTile;
  cudaError_t e;
  switch (desc_bytes) {
    case 48: e = launch_knn_words<12>(q, nq, t, nt, kr, train_index_offset, keys, part, splits, rows_per_split > 0 ? rows_per_split : kKnnTile, stream); break;
    case 64: e = launch_knn_words<16>(q, nq, t, nt, kr, train_index_offset, keys, part, splits, rows_per_split > 0 ? rows_per_split : kKnnTile, stream); break;
    case 128: e = launch_knn_words<32>(q, nq, t, nt, kr, train_index_offset, keys, part, splits, rows_per_split > 0 ? rows_per_split : kKnnTile, stream); break;
    default: return cudaErrorInvalidValue;
  }
  if (e != cudaSuccess) return e;
  if (splits > 1) {
    knn_merge_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, stream>>>(part, splits, nq, kr, keys);
    e = cudaGetLastError();
  }
  return e;
}

cudaError_t launch_knn_merge(const unsigned long long* gathered, int n_shards, long long nq, int k, unsigned long long* out,
                             cudaStream_t stream) {
  if (nq <= 0) return cudaSuccess;
  knn_merge_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, stream>>>(gathered, n_shards, nq, k, out);
  return cudaGetLastError();
}

cudaError_t launch_knn_unpack(const unsigned long long* keys, long long n, int32_t* idx, int32_t* dist, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  knn_unpack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(keys, n, idx, dist);
  return cudaGetLastError();
}

// brisk::Hamming::operator()(a, b, size) on n descriptor pairs (reference hamming.h:101-113, hamming-inl.h:85-134):
// popcount of the XOR over size / 16 whole 128-bit words -- bytes beyond the last whole word are not read.
// One warp per pair, one 32-bit word per lane and step.
__global__ void __launch_bounds__(256)
hamming_pairs_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, long long n, int desc_bytes, int32_t* __restrict__ dist) {
  const long long pair = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (pair >= n) return;
  const int words = (desc_bytes / 16) * 4;  // 32-bit words inside whole 128-bit words
  const uint8_t* pa = a + pair * desc_bytes;
  const uint8_t* pb = b + pair * desc_bytes;
  const bool aligned = (((uintptr_t)pa | (uintptr_t)pb) & 3) == 0;
  int d = 0;
  for (int w = lane; w < words; w += 32) {
    uint32_t x, y;
    if (aligned) { x = reinterpret_cast<const uint32_t*>(pa)[w]; y = reinterpret_cast<const uint32_t*>(pb)[w]; }
    else {
      x = pa[4 * w] | (pa[4 * w + 1] << 8) | (pa[4 * w + 2] << 16) | ((uint32_t)pa[4 * w + 3] << 24);
      y = pb[4 * w] | (pb[4 * w + 1] << 8) | (pb[4 * w + 2] << 16) | ((uint32_t)pb[4 * w + 3] << 24);
    }
    d += __popc(x ^ y);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
  if (lane == 0) dist[pair] = d;
}

cudaError_t launch_hamming_pairs(const uint8_t* a, const uint8_t* b, long long n, int desc_bytes, int32_t* dist, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  hamming_pairs_kernel<<<(unsigned)((n + 7) / 8), 256, 0, stream>>>(a, b, n, desc_bytes, dist);
  return cudaGetLastError();
}

}  // namespace briskb200
