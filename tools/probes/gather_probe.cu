// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/gather_probe tools/probes/gather_probe.cu
// Micro-benchmark: throughput of scattered 16-byte gathers from a block image (the access pattern of the
// descriptor sampler: one warp per key point, lanes at scattered offsets inside a window around it) through
// the load/store path (LDG.128) and through the texture path (tex1Dfetch<int4> / tex2D<int4>).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int W = 1920, H = 1080;

__device__ __forceinline__ uint32_t hash(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

template <int MODE>
__global__ void __launch_bounds__(256) gather(const int4* __restrict__ blocks, cudaTextureObject_t t1, cudaTextureObject_t t2, int n_kp, int radius,
                                              int rounds, int* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int kp = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (kp >= n_kp) return;
  const uint32_t hk = hash(kp * 2654435761u + 17);
  const int cx = radius + 2 + hk % (W - 2 * radius - 4), cy = radius + 2 + (hk >> 12) % (H - 2 * radius - 4);
  int acc = 0;
  for (int r = 0; r < rounds; ++r) {
    const uint32_t h = hash(hk + r * 32 + lane);
    const int x = cx - radius + (int)(h % (2 * radius + 1)), y = cy - radius + (int)((h >> 10) % (2 * radius + 1));
    int4 v;
    if (MODE == 0) v = __ldg(blocks + (long long)y * W + x);
    else if (MODE == 1) v = tex1Dfetch<int4>(t1, y * W + x);
    else v = tex2D<int4>(t2, (float)x, (float)y);
    acc += v.x ^ v.y ^ v.z ^ v.w;
  }
  if (acc == 0x12345678) out[kp] = acc;
}

int main() {
  const size_t n = (size_t)W * H;
  int4* d; CK(cudaMalloc(&d, n * 16));
  std::vector<int4> h(n);
  for (size_t i = 0; i < n; ++i) h[i] = make_int4((int)i, (int)(i * 3), (int)(i * 7), (int)(i * 11));
  CK(cudaMemcpy(d, h.data(), n * 16, cudaMemcpyHostToDevice));
  int* out; CK(cudaMalloc(&out, 1 << 24));
  cudaResourceDesc rd = {}; cudaTextureDesc td = {};
  rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = d; rd.res.linear.desc = cudaCreateChannelDesc<int4>(); rd.res.linear.sizeInBytes = n * 16;
  td.readMode = cudaReadModeElementType; td.filterMode = cudaFilterModePoint; td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
  cudaTextureObject_t t1, t2; CK(cudaCreateTextureObject(&t1, &rd, &td, nullptr));
  cudaResourceDesc r2 = {};
  r2.resType = cudaResourceTypePitch2D; r2.res.pitch2D.devPtr = d; r2.res.pitch2D.desc = cudaCreateChannelDesc<int4>();
  r2.res.pitch2D.width = W; r2.res.pitch2D.height = H; r2.res.pitch2D.pitchInBytes = (size_t)W * 16;
  CK(cudaCreateTextureObject(&t2, &r2, &td, nullptr));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int n_kp = 1 << 20, rounds = 16;
  for (int radius : {12, 30, 80}) {
    for (int mode = 0; mode < 3; ++mode) {
      float best = 1e9f;
      for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0));
        if (mode == 0) gather<0><<<n_kp / 8, 256>>>(d, t1, t2, n_kp, radius, rounds, out);
        if (mode == 1) gather<1><<<n_kp / 8, 256>>>(d, t1, t2, n_kp, radius, rounds, out);
        if (mode == 2) gather<2><<<n_kp / 8, 256>>>(d, t1, t2, n_kp, radius, rounds, out);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
      }
      const double fetches = (double)n_kp * rounds * 32;
      printf("radius %3d  %-12s %8.3f ms  %7.2f Gfetch/s  %6.2f cycles/warp-fetch/SM @1.9GHz\n", radius,
             mode == 0 ? "ldg.128" : mode == 1 ? "tex1Dfetch" : "tex2D", best, fetches / best * 1e-6,
             best * 1e-3 * 1.9e9 * 148 / ((double)n_kp * rounds));
    }
  }
  return 0;
}
