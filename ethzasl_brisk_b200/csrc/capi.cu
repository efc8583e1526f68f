// C ABI of libbrisk_b200.so (see include/brisk_b200.h): contexts, detector /
// extractor objects, chunked batch orchestration on one CUDA stream.
#include "../../include/brisk_b200.h"

#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "describe_logic.cuh"
#include "gcc_sort.cuh"
#include "harris_logic.cuh"
#include "kernels.h"
#include "pattern.h"

using namespace briskb200;

static_assert(sizeof(brisk_keypoint) == 28 && sizeof(KeyPoint) == 28, "key point layout");

namespace {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  // for buffers whose rows are copied out beyond what a kernel wrote (the strided result copies): never hand
  // uninitialised device memory to the caller
  cudaError_t ensure_zeroed(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    cudaError_t e = ensure(bytes);
    if (e == cudaSuccess) e = cudaMemset(p, 0, cap);
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

}  // namespace

// Per-chunk workspace + stream: two slots let the copies of one chunk overlap the kernels of the other.
struct Slot {
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[BRISK_STAGE_COUNT + 1] = {};
  cudaEvent_t done = nullptr;
  cudaEvent_t computed = nullptr;  // recorded after the last kernel of a chunk
  cudaStream_t side = nullptr;     // the integral image of a chunk is built here while the detector's kernels run
  cudaEvent_t fork = nullptr, join = nullptr;
  int32_t* h_counts = nullptr;  // pinned: counts of the chunk + [cap_counts] error flag
  size_t h_counts_cap = 0;
  DevBuf pyr, cm, bm, rowcnt, layer_start, corners, fwin, checks, kp_tmp, kp_valid, integral, rounds, surv;
  DevBuf tight, kps, kps_scratch, scales, counts, desc, masks, flag;
  DevBuf h_scores, h_pts, h_keep, h_sorted, h_layer_kept, h_occ, h_surv, h_layer_surv;
  DevBuf prov_base;  // ComputeScale: first key-point slot of every layer, [frame][kMaxLayers + 1]
  DevBuf* all[30] = {&pyr, &cm, &bm, &rowcnt, &layer_start, &corners, &fwin, &checks, &kp_tmp, &kp_valid, &integral, &rounds, &surv,
                     &tight, &kps, &kps_scratch, &scales, &counts, &desc, &masks, &flag,
                     &h_scores, &h_pts, &h_keep, &h_sorted, &h_layer_kept, &h_occ, &h_surv, &h_layer_surv, &prov_base};
};

struct brisk_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;  // caller-visible stream: work of a call is ordered after it
  bool own_stream = false;
  std::string err;
  size_t ws_limit = (size_t)8 << 30;
  bool timing = false;
  bool pipelining = true;
  int knn_variant = 3;  // 0: POPC kernel always; where they apply (k == 2, 48/64-byte rows) 1: mma.sync IMMA kernel, 2: tcgen05 kind::i8 kernel,
                        // 3 (default): tcgen05 kind::mxf4 (FP4) kernel
  float ms[BRISK_STAGE_COUNT] = {};
  int64_t launches = 0;
  int64_t raw_corners = 0;  // AGAST corners before NMS, summed over the frames of the last call (timing mode only)
  cudaEvent_t ev[2] = {};
  cudaEvent_t entry = nullptr;
  PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
  Slot slots[2];
  // brisk_detect_describe_async: the last chunk of the previous asynchronous call, still computing / not yet copied
  // back, and the status of the asynchronous calls since the last brisk_sync
  struct CallStatus { bool truncated = false, corner_overflow = false, internal_error = false, low_score = false, occupancy_oob = false; };
  struct Pending {
    int f0 = 0, c = 0;
    bool active = false, corners = false;
    KeyPoint* d_kps = nullptr; uint8_t* d_desc = nullptr;
    // where the results go and how the slot's pinned count block is laid out
    brisk_keypoint* kps = nullptr; int32_t* counts = nullptr; uint8_t* desc = nullptr;
    int cap = 0, desc_bytes = 0, chunk = 0, corner_cap = 0;
    bool kps_dev = false, counts_dev = false, desc_dev = false;
    CallStatus* st = nullptr;
  };
  struct AsyncKey {
    const void *det = nullptr, *ext = nullptr, *masks = nullptr;
    int n = 0, w = 0, h = 0, cap = 0; size_t stride = 0, frame_pitch = 0;
    bool imgs_dev = false, kps_dev = false, counts_dev = false, desc_dev = false;
    bool operator==(const AsyncKey& o) const {
      return det == o.det && ext == o.ext && (masks != nullptr) == (o.masks != nullptr) && n == o.n && w == o.w && h == o.h && cap == o.cap &&
             stride == o.stride && frame_pitch == o.frame_pitch && imgs_dev == o.imgs_dev && kps_dev == o.kps_dev &&
             counts_dev == o.counts_dev && desc_dev == o.desc_dev;
    }
  };
  Pending deferred;
  int deferred_slot = 0;
  AsyncKey deferred_key;
  CallStatus async_status;
  DevBuf knn_q, knn_t, knn_qx, knn_tx, knn_keys, knn_part, knn_idx, knn_dist, knn_mask, rad_counts, rad_offsets, rad_matches;
};

struct brisk_detector {
  brisk_ctx* ctx;
  int thresh, octaves, suppress;
  int corner_cap;  // 0 = auto
  int harris = 0;  // 1: Harris scale-space detector, 2: legacy single-scale HarrisFeatureDetector
  double radius = 0, abs_thr = 0;
  long long max_kpt = -1;
};

struct brisk_extractor {
  brisk_ctx* ctx;
  int device = 0;  // (kept here: the context may be gone by the time the extractor is destroyed)
  PatternHost host;
  DevBuf points, size_list, short_pairs, long_pairs, breaks, consts;
  PatternDev dev;
};

namespace {

int fail(brisk_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return code;
}

#define CU_OK(call)                                                                         \
  do {                                                                                      \
    cudaError_t e__ = (call);                                                               \
    if (e__ != cudaSuccess) {                                                               \
      cudaGetLastError();                                                                   \
      return fail(ctx, BRISK_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    }                                                                                       \
  } while (0)

bool is_device_ptr(const void* p) {
  if (!p) return false;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

int align_up(int v, int a) { return (v + a - 1) / a * a; }

// Layer sizes / scales of BriskScaleSpace::ConstructPyramid (reference
// brisk-scale-space.cc:64-90) and BriskLayer's constructors (brisk-layer.cc:53-95).
void build_geom(int w, int h, int octaves, PyramidGeom* g) {
  memset(g, 0, sizeof(*g));
  g->n_layers = octaves == 0 ? 1 : 2 * octaves;
  g->w0 = w; g->h0 = h;
  long long off = 0;
  for (int i = 0; i < g->n_layers; ++i) {
    LayerGeom& L = g->L[i];
    if (i == 0) { L.w = w; L.h = h; L.scale = 1.0f; L.offset = 0.0f; }
    else if (i == 1) { L.w = 2 * (w / 3); L.h = 2 * (h / 3); L.scale = (float)(g->L[0].scale * 1.5); L.offset = (float)(0.5 * L.scale - 0.5); }
    else { L.w = g->L[i - 2].w / 2; L.h = g->L[i - 2].h / 2; L.scale = g->L[i - 2].scale * 2; L.offset = (float)(0.5 * L.scale - 0.5); }
    L.pitch = align_up(std::max(L.w, 1), 16);
    L.off = off;
    off += (long long)L.pitch * std::max(L.h, 1);
    off = (off + 255) / 256 * 256;
  }
  g->frame_elems = off;
}

struct Timer {
  brisk_ctx* ctx;
  Slot* slot;
  Timer(brisk_ctx* c, Slot* s) : ctx(c), slot(s) {}
  void mark(int i) { if (ctx->timing) cudaEventRecord(slot->ev[i], slot->stream); }
};

int encode_map(brisk_ctx* ctx, const void* base, int w, int h, int n, size_t pitch, size_t frame_stride, CUtensorMap* map) {
  cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)frame_stride};
  cuuint32_t box[3] = {192, 96, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = ctx->encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ctx, BRISK_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
  return BRISK_OK;
}

struct Plan {
  PyramidGeom g;
  HarrisWorkspace hw;  // occupancy geometry (Harris only)
  DetectWorkspace ws;  // geometry part only; pointers come from slot_ws()
  int chunk;
  int n_slots;
  size_t integral_elems;  // per frame
};

// host_io: the frames come from host memory; any_host: some buffer of the call lives in host memory.  With
// everything resident on the device there is nothing to overlap and one slot with large chunks is faster
// (measured: 92 vs 105 ms per 1024 1080p frames), so the two-slot pipeline is only used when copies exist.
int drain_async(brisk_ctx* ctx);

int make_plan(brisk_ctx* ctx, const brisk_detector* det, const brisk_extractor* ext, int n, int w, int h, int cap, Plan* plan,
              bool host_io = false, bool any_host = true, bool two_slots = false) {
  // the slots' buffers are about to be (re)used: whatever an asynchronous call left in flight is completed first
  // (run_batch takes a chained call's pending chunk out of the context before it plans)
  if (ctx->deferred.active) { const int rc = drain_async(ctx); if (rc) return rc; }
  const bool pipelining = ctx->pipelining && any_host;
  const int octaves = det ? det->octaves : 0;
  build_geom(w, h, octaves, &plan->g);
  const PyramidGeom& g = plan->g;
  DetectWorkspace& ws = plan->ws;
  memset(&ws, 0, sizeof(ws));
  ws.total_rows = 0;
  for (int i = 0; i < g.n_layers; ++i) { ws.row_off[i] = ws.total_rows; ws.total_rows += g.L[i].h; }
  ws.row_off[g.n_layers] = ws.total_rows;
  int ccap = det ? det->corner_cap : 0;
  const bool harris = det && det->harris;
  if (ccap <= 0) ccap = std::min(std::max((int)(((long long)w * h) / (harris ? 10 : 24)), harris ? 8192 : 4096), 1 << 20);
  memset(&plan->hw, 0, sizeof(plan->hw));
  if (harris && det->harris == 2) {
    // harris-feature-detector.cc:313-315: one half-resolution map (+ slack for the clearing stores)
    plan->hw.occ_h[0] = h / 2 + 32; plan->hw.occ_w[0] = w / 2 + 32;
    plan->hw.occ_frame_bytes = ((long long)plan->hw.occ_h[0] * plan->hw.occ_w[0] + 64 + 255) / 256 * 256;
  } else if (harris && det->radius > 0.0) {
    // occupancy maps of EnforceKeyPointUniformity (uniformity-enforcement-inl.h:62-66); none for key-point bucketing
    const float scaling = (float)(15.0 / (double)(float)(det->radius == 0 ? 1.0 : det->radius));
    long long off = 0;
    for (int i = 0; i < g.n_layers; ++i) {
      plan->hw.occ_h[i] = (int)(g.L[i].h * std::ceil((double)scaling) + 32);
      plan->hw.occ_w[i] = (int)(g.L[i].w * std::ceil((double)scaling) + 32);
      plan->hw.occ_off[i] = off;
      off += ((long long)plan->hw.occ_h[i] * plan->hw.occ_w[i] + 64 + 255) / 256 * 256;
    }
    plan->hw.occ_frame_bytes = off;
  }
  ws.corner_cap = det ? ccap : 0;
  plan->integral_elems = ext ? (size_t)w * h * 4 : 0;  // one 2x2 block of the integral image per pixel
  const size_t integral_aux = ext ? (size_t)integral_aux_elems(w, h) : 0;
  const int desc_bytes = ext ? ext->dev.desc_bytes : 0;
  size_t per_frame = (size_t)g.frame_elems;  // image planes
  if (det && !harris) per_frame += (size_t)g.frame_elems * 3 + (size_t)ws.total_rows * 4 + (size_t)ws.corner_cap * (4 + 32 + 32 + 28 + 1 + 4);
  if (harris) per_frame += (size_t)g.frame_elems * 4 + (size_t)ws.total_rows * 4 + (size_t)ws.corner_cap * 25 + (size_t)plan->hw.occ_frame_bytes;
  per_frame += (plan->integral_elems + integral_aux) * 4 + (size_t)cap * (28 * 2 + 4 + desc_bytes) + 2 * (size_t)w * h /* mask, tight copy */;
  // two slots share the workspace limit; at least two chunks when there is more than one frame, so
  // that copies and the serial tail of one chunk overlap the kernels of the other
  long long max_chunk = (long long)(ctx->ws_limit / 2 / std::max<size_t>(per_frame, 1));
  if (max_chunk < 1) max_chunk = 1;
  if (max_chunk > 32768) max_chunk = 32768;
  if (!pipelining) max_chunk = std::min<long long>(max_chunk * 2, 32768);  // one slot gets the whole budget
  long long n_chunks = (n + max_chunk - 1) / max_chunk;
  if (n_chunks < 2 && n > 1 && pipelining) n_chunks = 2;
  // host buffers: the first chunk's upload and the last chunk's download are not hidden behind any kernel, so cut
  // the batch finer (down to about 150 MB of pixels per chunk) -- but into at most 4 chunks, and none below 256
  // frames when the batch allows it: the per-frame kernels (tie chain, compaction, border cull: one CTA per frame,
  // 296 resident CTAs of the chain kernel) run a 128-frame chunk no faster than a 296-frame one (measured: 8 chunks
  // of 128 1080p frames cost 12 ms per 1024 frames more than 4 of 256)
  if (host_io && pipelining) {
    const long long by_size = std::max<long long>(1, (long long)n * w * h / (150ll << 20));
    const long long by_frames = std::max<long long>(2, n / 256);
    n_chunks = std::max(n_chunks, std::min<long long>(std::min<long long>(4, by_frames), by_size));
  }
  if (n_chunks < 1) n_chunks = 1;
  const long long chunk = std::max<long long>(1, (n + n_chunks - 1) / n_chunks);  // equal-sized chunks, no tiny tail
  plan->chunk = (int)chunk;
  plan->n_slots = ((n > plan->chunk || two_slots) && pipelining) ? 2 : 1;  // (asynchronous calls alternate the slots from call to call)
  const size_t c = (size_t)plan->chunk;
  for (int si = 0; si < plan->n_slots; ++si) {
    Slot& sl = ctx->slots[si];
    CU_OK(sl.pyr.ensure(c * g.frame_elems));
    if (harris) {
      CU_OK(sl.rowcnt.ensure(c * ws.total_rows * 4));
      CU_OK(sl.layer_start.ensure(c * (kMaxLayers + 1) * 4));
      CU_OK(sl.h_scores.ensure(c * g.frame_elems * 4));
      CU_OK(sl.h_pts.ensure(c * ws.corner_cap * 8));
      CU_OK(sl.h_keep.ensure(c * ws.corner_cap));
      CU_OK(sl.h_sorted.ensure(c * ws.corner_cap * 8));
      CU_OK(sl.h_surv.ensure(c * ws.corner_cap * 8));
      CU_OK(sl.h_layer_kept.ensure(c * kMaxLayers * 4));
      CU_OK(sl.h_layer_surv.ensure(c * kMaxLayers * 4));
      CU_OK(sl.h_occ.ensure(c * (size_t)plan->hw.occ_frame_bytes));
    } else if (det) {
      CU_OK(sl.cm.ensure(c * g.frame_elems * 2));
      CU_OK(sl.bm.ensure(c * g.frame_elems));
      CU_OK(sl.rowcnt.ensure(c * ws.total_rows * 4));
      CU_OK(sl.layer_start.ensure(c * (kMaxLayers + 1) * 4));
      CU_OK(sl.corners.ensure(c * ws.corner_cap * 4));
      CU_OK(sl.fwin.ensure(c * ws.corner_cap * 32));
      CU_OK(sl.checks.ensure(c * ws.corner_cap * 32));
      CU_OK(sl.kp_tmp.ensure(c * ws.corner_cap * 28));
      CU_OK(sl.kp_valid.ensure(c * ws.corner_cap));
      CU_OK(sl.rounds.ensure(c * kTieStride * 4));
      CU_OK(sl.surv.ensure(c * (size_t)ws.corner_cap * 4));
    }
    if (ext) {
      CU_OK(sl.integral.ensure(c * (plan->integral_elems + integral_aux) * 4));
      CU_OK(sl.kps_scratch.ensure(c * cap * 28));
      CU_OK(sl.scales.ensure(c * cap * 4));
    }
    CU_OK(sl.flag.ensure(16));
    if (sl.h_counts_cap < 2 * c + 4) {  // counts [c], flags [4], raw corner totals [c] (timing mode)
      if (sl.h_counts) cudaFreeHost(sl.h_counts);
      sl.h_counts = nullptr; sl.h_counts_cap = 0;
      CU_OK(cudaMallocHost(&sl.h_counts, (2 * c + 4) * sizeof(int32_t)));
      sl.h_counts_cap = 2 * c + 4;
    }
  }
  return BRISK_OK;
}

// Device pointers of a slot's workspace for the kernels.
DetectWorkspace slot_ws(const Plan& plan, const Slot& sl) {
  DetectWorkspace ws = plan.ws;
  ws.pyr = sl.pyr.as<uint8_t>(); ws.cm = sl.cm.as<uint16_t>(); ws.bm = sl.bm.as<uint8_t>();
  ws.rowcnt = sl.rowcnt.as<int>(); ws.layer_start = sl.layer_start.as<int>(); ws.corners = sl.corners.as<uint32_t>();
  ws.fwin = sl.fwin.as<uint8_t>(); ws.checks = sl.checks.as<float>(); ws.kp_tmp = sl.kp_tmp.as<KeyPoint>();
  ws.kp_valid = sl.kp_valid.as<uint8_t>(); ws.n_ties = sl.rounds.as<int>(); ws.surv = sl.surv.as<int>();
  return ws;
}

// Bring `count` frames starting at `imgs` into the pyramid block (layer 0 of
// every frame) or, when they already sit in device memory with TMA-compatible
// alignment, describe them in place.  Returns the tensor map the pyramid
// kernel reads and whether it has to materialise layer 0 in the block.
int stage_input(brisk_ctx* ctx, Slot& sl, const Plan& plan, const uint8_t* imgs, int count, int w, int h, size_t stride,
                size_t frame_pitch, CUtensorMap* map, int* write_l0, bool* staged_tight = nullptr) {
  const PyramidGeom& g = plan.g;
  const bool dev = is_device_ptr(imgs);
  const bool aligned = dev && ((uintptr_t)imgs % 16 == 0) && (stride % 16 == 0) && (frame_pitch % 16 == 0);
  if (staged_tight) *staged_tight = false;
  if (aligned) {
    *write_l0 = 1;
    return encode_map(ctx, imgs, w, h, count, stride, frame_pitch, map);
  }
  // Host frames that are tightly packed and contiguous travel in ONE copy into the staging buffer, which
  // the pyramid kernel then reads like a caller's device batch (and the sampler reuses as its packed image).
  if (staged_tight && !dev && w % 16 == 0 && stride == (size_t)w && (count == 1 || frame_pitch == (size_t)w * h)) {
    CU_OK(cudaMemcpyAsync(sl.tight.p, imgs, (size_t)count * w * h, cudaMemcpyHostToDevice, sl.stream));
    *write_l0 = 1;
    *staged_tight = true;
    return encode_map(ctx, sl.tight.as<uint8_t>(), w, h, count, (size_t)w, (size_t)w * h, map);
  }
  for (int f = 0; f < count; ++f)
    CU_OK(cudaMemcpy2DAsync(sl.pyr.as<uint8_t>() + (size_t)f * g.frame_elems + g.L[0].off, g.L[0].pitch,
                            imgs + (size_t)f * frame_pitch, stride, w, h, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                            sl.stream));
  *write_l0 = 0;
  return encode_map(ctx, sl.pyr.as<uint8_t>() + g.L[0].off, w, h, count, g.L[0].pitch, (size_t)g.frame_elems, map);
}

// The occupancy map of EnforceKeyPointUniformity (uniformity-enforcement-inl.h:62-66) takes ceil(15 / radius)^2 bytes
// per pixel; the kernel indexes one layer's map with an int.
const char* const kOccupancyTooLarge = "uniformityRadius is too small for this image size: the occupancy map of layer 0 would exceed 2 GiB";
bool occupancy_fits(double radius, int w, int h) {
  const float scaling = (float)(15.0 / (double)(float)(radius == 0 ? 1.0 : radius));
  const double c = std::ceil((double)scaling);
  return ((double)h * c + 32.0) * ((double)w * c + 32.0) + 64.0 < 2147483648.0;
}

int check_image_args(brisk_ctx* ctx, const uint8_t* imgs, int n, int w, int h, size_t stride, size_t frame_pitch) {
  if (!ctx) return BRISK_ERR_INVALID;
  if (!imgs || n < 0 || w <= 0 || h <= 0) return fail(ctx, BRISK_ERR_INVALID, "bad image arguments");
  if (stride < (size_t)w || (n > 1 && frame_pitch < stride * (size_t)h)) return fail(ctx, BRISK_ERR_INVALID, "bad image strides");
  if (w > 8191 || h > 8191) return fail(ctx, BRISK_ERR_UNSUPPORTED, "images larger than 8191 pixels per side are not supported");
  return BRISK_OK;
}

// Key-point lists in, key-point lists out (ComputeScale, passed key points): validation and staging of the caller's
// host or device buffers, chunk by chunk on one slot.
struct KpLists {
  const brisk_keypoint* kps_in; const int32_t* counts_in; int cap_in;
  brisk_keypoint* kps_out; int32_t* counts_out; int cap_out;
  bool in_dev = false, cin_dev = false, out_dev = false, cout_dev = false, truncated = false;
  std::vector<int32_t> h_cin;  // counts_in on the host
  int in_max = 0;              // longest input list
  const KeyPoint* d_in = nullptr; const int* d_cin = nullptr; KeyPoint* d_out = nullptr; int* d_cout = nullptr;  // current chunk

  int init(brisk_ctx* ctx, int n, const char* empty_msg) {
    if (!kps_in || !counts_in || cap_in <= 0 || !kps_out || !counts_out || cap_out <= 0) return fail(ctx, BRISK_ERR_INVALID, "bad key point arguments");
    in_dev = is_device_ptr(kps_in); cin_dev = is_device_ptr(counts_in);
    out_dev = is_device_ptr(kps_out); cout_dev = is_device_ptr(counts_out);
    h_cin.resize(n);
    if (cin_dev) {
      CU_OK(cudaStreamSynchronize(ctx->stream));
      CU_OK(cudaMemcpy(h_cin.data(), counts_in, (size_t)n * 4, cudaMemcpyDeviceToHost));
    } else if (n) memcpy(h_cin.data(), counts_in, (size_t)n * 4);
    for (int f = 0; f < n; ++f) {
      if (h_cin[f] < 0 || h_cin[f] > cap_in) return fail(ctx, BRISK_ERR_INVALID, "counts_in must be in [0, cap_in]");
      if (h_cin[f] == 0) return fail(ctx, BRISK_ERR_UNSUPPORTED, empty_msg);
      in_max = std::max(in_max, h_cin[f]);
    }
    return BRISK_OK;
  }
  int reserve(brisk_ctx* ctx, Slot& sl, size_t c_max) {
    if (!in_dev) CU_OK(sl.kps_scratch.ensure(c_max * cap_in * 28));
    if (!cin_dev) CU_OK(sl.scales.ensure(c_max * 4));
    if (!out_dev) CU_OK(sl.kps.ensure(c_max * cap_out * 28));
    if (!cout_dev) CU_OK(sl.counts.ensure(c_max * 4));
    return BRISK_OK;
  }
  // device views of frames [f0, f0 + c); host lists are uploaded (only their used part)
  int upload(brisk_ctx* ctx, Slot& sl, int f0, int c) {
    d_in = reinterpret_cast<const KeyPoint*>(kps_in) + (size_t)f0 * cap_in;
    d_cin = counts_in + f0;
    if (!in_dev) {
      for (int f = 0; f < c; ++f)
        CU_OK(cudaMemcpyAsync(sl.kps_scratch.as<KeyPoint>() + (size_t)f * cap_in, kps_in + (size_t)(f0 + f) * cap_in,
                              (size_t)h_cin[f0 + f] * 28, cudaMemcpyHostToDevice, sl.stream));
      d_in = sl.kps_scratch.as<KeyPoint>();
    }
    if (!cin_dev) {
      CU_OK(cudaMemcpyAsync(sl.scales.p, h_cin.data() + f0, (size_t)c * 4, cudaMemcpyHostToDevice, sl.stream));
      d_cin = sl.scales.as<int>();
    }
    d_out = out_dev ? reinterpret_cast<KeyPoint*>(kps_out) + (size_t)f0 * cap_out : sl.kps.as<KeyPoint>();
    d_cout = cout_dev ? counts_out + f0 : sl.counts.as<int>();
    return BRISK_OK;
  }
  // waits for the chunk, returns its error flag, copies counts and the produced rows back
  int download(brisk_ctx* ctx, Slot& sl, int chunk, int f0, int c, int* flag) {
    CU_OK(cudaMemcpyAsync(sl.h_counts, d_cout, (size_t)c * 4, cudaMemcpyDeviceToHost, sl.stream));
    CU_OK(cudaMemcpyAsync(sl.h_counts + chunk, sl.flag.p, 4, cudaMemcpyDeviceToHost, sl.stream));
    CU_OK(cudaStreamSynchronize(sl.stream));
    *flag = sl.h_counts[chunk];
    if (!cout_dev) memcpy(counts_out + f0, sl.h_counts, (size_t)c * 4);
    for (int f = 0; f < c; ++f) {
      const int m = std::min(sl.h_counts[f], cap_out);
      if (sl.h_counts[f] > cap_out) truncated = true;
      if (m > 0 && !out_dev)
        CU_OK(cudaMemcpyAsync(kps_out + (size_t)(f0 + f) * cap_out, d_out + (size_t)f * cap_out, (size_t)m * 28, cudaMemcpyDeviceToHost, sl.stream));
    }
    CU_OK(cudaStreamSynchronize(sl.stream));
    return BRISK_OK;
  }
};

// Core batch driver: detect and/or describe, chunk by chunk.
// Second half of a chunk: wait for its kernels, then copy exactly the produced rows back (queued on the slot's stream).
int finish_chunk(brisk_ctx* ctx, int si, brisk_ctx::Pending& pd) {
  if (!pd.active) return BRISK_OK;
  Slot& sl = ctx->slots[si];
  pd.active = false;
  CU_OK(cudaStreamSynchronize(sl.stream));
  brisk_ctx::CallStatus& st = *pd.st;
  const int32_t flag = sl.h_counts[pd.chunk];
  if (sl.h_counts[pd.chunk + 1] == 5) st.low_score = true;
  else if (flag == 6) st.occupancy_oob = true;
  else if (flag == 2) st.internal_error = true;
  else if (flag) st.corner_overflow = true;
  if (!pd.counts_dev) memcpy(pd.counts + pd.f0, sl.h_counts, (size_t)pd.c * 4);
  if (pd.corners)
    for (int f = 0; f < pd.c; ++f) ctx->raw_corners += std::min(sl.h_counts[pd.chunk + 4 + f], pd.corner_cap);
  Timer tm(ctx, &sl);
  tm.mark(7);
  // The produced rows of the whole chunk go back in ONE strided copy per array (rows = frames, width = the longest
  // list of the chunk): a few per cent more bytes than frame-by-frame copies of the exact lengths, but 2 DMA
  // descriptors per chunk instead of 2 per frame (2 x 1024 driver calls per step on the benched workload, per rank).
  const int cap = pd.cap;
  int longest = 0;
  for (int f = 0; f < pd.c; ++f) {
    if (sl.h_counts[f] > cap) st.truncated = true;
    longest = std::max(longest, std::min(sl.h_counts[f], cap));
  }
  if (longest > 0) {
    if (!pd.kps_dev)
      CU_OK(cudaMemcpy2DAsync(pd.kps + (size_t)pd.f0 * cap, (size_t)cap * 28, pd.d_kps, (size_t)cap * 28, (size_t)longest * 28, (size_t)pd.c,
                              cudaMemcpyDeviceToHost, sl.stream));
    if (pd.desc && !pd.desc_dev)
      CU_OK(cudaMemcpy2DAsync(pd.desc + (size_t)pd.f0 * cap * pd.desc_bytes, (size_t)cap * pd.desc_bytes, pd.d_desc, (size_t)cap * pd.desc_bytes,
                              (size_t)longest * pd.desc_bytes, (size_t)pd.c, cudaMemcpyDeviceToHost, sl.stream));
  }
  tm.mark(8);
  // the asynchronous call waits for exactly these copies before it returns (not for whatever is queued behind them)
  CU_OK(cudaEventRecord(sl.done, sl.stream));
  return BRISK_OK;
}

int report_status(brisk_ctx* ctx, const brisk_ctx::CallStatus& st) {
  if (st.low_score)
    return fail(ctx, BRISK_ERR_UNSUPPORTED, "a detected corner scores <= 2 (possible for thresh < 20 only): the reference's score cache does not keep such scores "
                                            "and its result becomes order dependent; not supported");
  if (st.occupancy_oob)
    return fail(ctx, BRISK_ERR_UNSUPPORTED, "HarrisFeatureDetector: a key point's occupancy-map indices leave the map -- the reference indexes it with x as "
                                            "the row (harris-feature-detector.cc:322-329) and accesses memory out of bounds for this image shape "
                                            "(landscape images); no defined result to match");
  if (st.internal_error) return fail(ctx, BRISK_ERR_CUDA, "internal error: tie resolution did not converge");
  if (st.corner_overflow) return fail(ctx, BRISK_ERR_CAPACITY, "raw corner capacity exceeded; raise it with brisk_detector_set_corner_capacity");
  if (st.truncated) return fail(ctx, BRISK_ERR_CAPACITY, "key point capacity (cap) exceeded; counts hold the true numbers");
  return BRISK_OK;
}

// Completes what brisk_detect_describe_async left in flight and reports the status of the asynchronous calls since the
// last time.  Every other entry point that uses the slots calls it first.
int drain_async(brisk_ctx* ctx) {
  if (ctx->deferred.active) {
    CU_OK(cudaSetDevice(ctx->device));
    const int si = ctx->deferred_slot;
    const int rc = finish_chunk(ctx, si, ctx->deferred);
    if (rc) return rc;
    CU_OK(cudaStreamSynchronize(ctx->slots[si].stream));
  }
  const brisk_ctx::CallStatus st = ctx->async_status;
  ctx->async_status = brisk_ctx::CallStatus();
  return report_status(ctx, st);
}

int run_batch(brisk_ctx* ctx, brisk_detector* det, brisk_extractor* ext, const uint8_t* imgs, int n, int w, int h,
              size_t stride, size_t frame_pitch, const uint8_t* masks, brisk_keypoint* kps, int32_t* counts, int cap,
              uint8_t* desc, bool async = false) {
  int rc = check_image_args(ctx, imgs, n, w, h, stride, frame_pitch);
  if (rc) return rc;
  if (!kps || !counts || cap <= 0 || (ext && !desc)) return fail(ctx, BRISK_ERR_INVALID, "bad output arguments");
  CU_OK(cudaSetDevice(ctx->device));
  // An asynchronous call chains onto the previous one when it has the same shape (same plan, same buffers): its first
  // chunk is queued BEFORE the previous call's last chunk is collected, so that upload runs under the previous call's
  // kernels and the previous download under this call's.  Anything else completes the pending work first.
  const bool any_host = !is_device_ptr(imgs) || !is_device_ptr(kps) || !is_device_ptr(counts) || (desc && !is_device_ptr(desc)) ||
                        (masks && !is_device_ptr(masks));
  if (ctx->timing || !ctx->pipelining || !any_host || n == 0) async = false;
  brisk_ctx::AsyncKey key;
  key.det = det; key.ext = ext; key.masks = masks; key.n = n; key.w = w; key.h = h; key.cap = cap; key.stride = stride; key.frame_pitch = frame_pitch;
  key.imgs_dev = is_device_ptr(imgs); key.kps_dev = is_device_ptr(kps); key.counts_dev = is_device_ptr(counts); key.desc_dev = desc && is_device_ptr(desc);
  const bool chained = async && ctx->deferred.active && key == ctx->deferred_key;
  if (!chained) { rc = drain_async(ctx); if (rc) return rc; }
  // the previous call's last chunk travels with this call from here on; put back if the call fails before touching it
  brisk_ctx::Pending carried;
  struct Restore {
    brisk_ctx* c; brisk_ctx::Pending* p;
    ~Restore() { if (p->active) c->deferred = *p; }
  } restore{ctx, &carried};
  if (chained) { carried = ctx->deferred; ctx->deferred.active = false; }
  memset(ctx->ms, 0, sizeof(ctx->ms));
  ctx->launches = 0;
  ctx->raw_corners = 0;
  if (n == 0) return BRISK_OK;
  if (det && det->harris == 2) {
    if (ext) return fail(ctx, BRISK_ERR_INVALID, "the legacy HarrisFeatureDetector has no fused extraction; call detect, then describe");
    if (!(det->radius > 0.0)) return fail(ctx, BRISK_ERR_INVALID, "HarrisFeatureDetector needs a positive radius");
    // narrower images never enter the SSE loops of GetCovarEntries (harris-feature-detector.cc:100): all scores stay 0
    if (w - 2 < 16 || h < 5) return fail(ctx, BRISK_ERR_UNSUPPORTED, "HarrisFeatureDetector needs at least 18 columns and 5 rows");
  } else if (det && det->harris) {
    if (det->octaves < 0 || 2 * det->octaves > kMaxLayers) return fail(ctx, BRISK_ERR_UNSUPPORTED, "octaves must be in [0, 6]");
    if (!(det->radius > 0.0)) {
      // KeyPointBucketing (key-point-bucketing-inl.h:74-112): the reference reserves maxNumKpt entries up front, so the default
      // maxNumKpt = SIZE_MAX throws std::length_error there; a finite limit is required, and 4 buckets per axis need > 4 pixels
      if (det->max_kpt <= 0) return fail(ctx, BRISK_ERR_UNSUPPORTED, "key point bucketing (uniformityRadius <= 0) needs a finite maxNumKpt > 0");
    } else if (!occupancy_fits(det->radius, w, h)) return fail(ctx, BRISK_ERR_UNSUPPORTED, kOccupancyTooLarge);
  } else if (det) {
    // The closed form of the lazy score cache (nms_logic.cuh) relies on every detected corner holding a score > 2 (such
    // cache entries are returned whatever threshold is asked, brisk-layer.cc:124-126).  A corner's score is its
    // threshold-map value T, and a corner needs 9 ring pixels more than b = (max(T, 10) * thresh) / 100 away from the
    // centre inside a disk whose range is T, i.e. T >= thresh / 10 + 1: that is > 2 for every thresh >= 20.
    // Below 20 the same holds unless some corner actually scores <= 2 (rare down to thresh 10, common below): checked on the
    // detected corners of every batch (launch_corner_score_check).
    if (det->thresh < 1 || det->thresh > 255) return fail(ctx, BRISK_ERR_INVALID, "AGAST threshold must be in [1, 255]");
    if (det->octaves < 0 || 2 * det->octaves > kMaxLayers) return fail(ctx, BRISK_ERR_UNSUPPORTED, "octaves must be in [0, 6]");
    // suppressScaleNonmaxima = false (brisk-scale-space.cc:131-170): with one layer it is the single-layer branch
    // (:172-209) word for word; with more layers the reference indexes layer 0's corner list with the counts of
    // layer i (`agastPoints.at(0)[n]`, :137), which reads past its end whenever a coarser layer has more corners
    // -- undefined behaviour, nothing to be bit-exact with.
    if (!det->suppress && det->octaves != 0)
      return fail(ctx, BRISK_ERR_UNSUPPORTED, "suppressScaleNonmaxima=false is only defined for octaves == 0 (the reference reads out of bounds otherwise)");
  }
  {
    PyramidGeom probe;
    build_geom(w, h, det ? det->octaves : 0, &probe);
    for (int i = 0; i < probe.n_layers; ++i)
      if (probe.L[i].w < 8 || probe.L[i].h < 8) return fail(ctx, BRISK_ERR_INVALID, "image too small: every pyramid layer must be at least 8x8");
  }
  Plan plan;
  rc = make_plan(ctx, det, ext, n, w, h, cap, &plan, !is_device_ptr(imgs), any_host, async);
  if (rc) return rc;
  const PyramidGeom& g = plan.g;
  const int desc_bytes = ext ? ext->dev.desc_bytes : 0;
  const bool kps_dev = is_device_ptr(kps), counts_dev = is_device_ptr(counts), desc_dev = desc && is_device_ptr(desc);
  const bool masks_dev = masks && is_device_ptr(masks);
  for (int si = 0; si < plan.n_slots; ++si) {
    Slot& sl = ctx->slots[si];
    if (!kps_dev) CU_OK(sl.kps.ensure_zeroed((size_t)plan.chunk * cap * 28));
    if (!counts_dev) CU_OK(sl.counts.ensure((size_t)plan.chunk * 4));
    if (ext && !desc_dev) CU_OK(sl.desc.ensure_zeroed((size_t)plan.chunk * cap * desc_bytes));
    if (det && masks && !masks_dev) CU_OK(sl.masks.ensure((size_t)plan.chunk * w * h));
    if ((ext && g.L[0].pitch != w) || !is_device_ptr(imgs)) CU_OK(sl.tight.ensure((size_t)plan.chunk * ((size_t)w * h + 64)));
  }
  // order the slots' streams after whatever the caller queued on the context stream
  CU_OK(cudaEventRecord(ctx->entry, ctx->stream));
  for (int si = 0; si < plan.n_slots; ++si) CU_OK(cudaStreamWaitEvent(ctx->slots[si].stream, ctx->entry, 0));

  brisk_ctx::CallStatus local_status;
  brisk_ctx::CallStatus* status = async ? &ctx->async_status : &local_status;
  brisk_ctx::Pending pend[2];
  if (chained) { pend[ctx->deferred_slot] = carried; carried.active = false; }
  auto finish = [&](int si) -> int { return finish_chunk(ctx, si, pend[si]); };
  auto collect_timing = [&](int si) {
    if (!ctx->timing) return;
    static const int stage_of[8] = {BRISK_STAGE_H2D, BRISK_STAGE_PYRAMID, BRISK_STAGE_DETECT, BRISK_STAGE_LISTS, BRISK_STAGE_NMS,
                                    BRISK_STAGE_INTEGRAL, BRISK_STAGE_DESCRIBE, BRISK_STAGE_D2H};
    Slot& sl = ctx->slots[si];
    cudaStreamSynchronize(sl.stream);
    for (int i = 0; i < 8; ++i) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, sl.ev[i], sl.ev[i + 1]) == cudaSuccess) ctx->ms[stage_of[i]] += ms; else cudaGetLastError();
    }
  };

  // a chained call starts on the slot the previous call's last chunk does not occupy
  const int first_index = chained ? (ctx->deferred_slot ^ 1) : 0;
  int chunk_index = first_index;
  // With host frames the upload of the first chunk is not hidden behind any kernel (unless the call is chained): start
  // with a quarter chunk.
  const int first_chunk = (!chained && plan.n_slots == 2 && !is_device_ptr(imgs) && n >= 4 * plan.chunk) ? std::max(1, plan.chunk / 4) : plan.chunk;
  int last_slot = 0;
  for (int f0 = 0, step = first_chunk; f0 < n; f0 += step, step = plan.chunk, ++chunk_index) {
    const int si = chunk_index % plan.n_slots;
    Slot& sl = ctx->slots[si];
    // the slot's previous chunk must be fully drained (results copied) before its buffers are reused
    if (pend[si].active) { rc = finish(si); if (rc) return rc; }
    if (chunk_index >= first_index + plan.n_slots) collect_timing(si);
    last_slot = si;
    const int c = std::min(step, n - f0);
    const DetectWorkspace ws = slot_ws(plan, sl);
    KeyPoint* d_kps = kps_dev ? reinterpret_cast<KeyPoint*>(kps) + (size_t)f0 * cap : sl.kps.as<KeyPoint>();
    int* d_counts = counts_dev ? counts + f0 : sl.counts.as<int>();
    uint8_t* d_desc = ext ? (desc_dev ? desc + (size_t)f0 * cap * desc_bytes : sl.desc.as<uint8_t>()) : nullptr;
    Timer tm(ctx, &sl);

    tm.mark(0);
    CUtensorMap map;
    int write_l0 = 0;
    bool staged_tight = false;
    rc = stage_input(ctx, sl, plan, imgs + (size_t)f0 * frame_pitch, c, w, h, stride, frame_pitch, &map, &write_l0, &staged_tight);
    if (rc) return rc;
    const uint8_t* d_masks = nullptr;
    long long mask_fs = 0; int mask_pitch = 0;
    if (det && masks) {
      if (masks_dev) { d_masks = masks + (size_t)f0 * frame_pitch; mask_fs = (long long)frame_pitch; mask_pitch = (int)stride; }
      else {
        for (int f = 0; f < c; ++f)
          CU_OK(cudaMemcpy2DAsync(sl.masks.as<uint8_t>() + (size_t)f * w * h, w, masks + (size_t)(f0 + f) * frame_pitch, stride, w, h,
                                  cudaMemcpyHostToDevice, sl.stream));
        d_masks = sl.masks.as<uint8_t>(); mask_fs = (long long)w * h; mask_pitch = w;
      }
    }
    if (!det) {
      // describe only: key points come from the caller
      if (!kps_dev) CU_OK(cudaMemcpyAsync(d_kps, kps + (size_t)f0 * cap, (size_t)c * cap * 28, cudaMemcpyHostToDevice, sl.stream));
      if (!counts_dev) CU_OK(cudaMemcpyAsync(d_counts, counts + f0, (size_t)c * 4, cudaMemcpyHostToDevice, sl.stream));
    }
    tm.mark(1);
    // The two slots overlap COPIES with kernels, not kernels with kernels: kernels of two chunks running side
    // by side slow each other down by more than the overlap gains (measured), so a chunk's kernels start once
    // the other slot's kernels are done, while its upload ran ahead of them and the other's download follows.
    if (plan.n_slots == 2 && (chunk_index > first_index || chained)) CU_OK(cudaStreamWaitEvent(sl.stream, ctx->slots[si ^ 1].computed, 0));
    if (det || write_l0) {
      CU_OK(launch_pyramid(map, g, ws.pyr, c, write_l0, sl.stream));
      ctx->launches += 1;
    }
    tm.mark(2);
    // The integral image only needs layer 0, so it could be built on a side stream while the detector's kernels run
    // (BRISK_B200_INTEGRAL_ASIDE=1).  That paid in round 1; with the present kernels it is 1 % slower for resident
    // batches and even end to end, so the default builds it in line, right before the descriptor kernel.
    static const bool aside_wanted = getenv("BRISK_B200_INTEGRAL_ASIDE") && atoi(getenv("BRISK_B200_INTEGRAL_ASIDE")) == 1;
    const bool integral_aside = aside_wanted && det && ext && !ctx->timing;
    if (integral_aside) {
      CU_OK(cudaEventRecord(sl.fork, sl.stream));
      CU_OK(cudaStreamWaitEvent(sl.side, sl.fork, 0));
      CU_OK(launch_integral(ws.pyr + g.L[0].off, g.frame_elems, g.L[0].pitch, w, h, c, sl.integral.as<int32_t>(), sl.side));
      CU_OK(cudaEventRecord(sl.join, sl.side));
      ctx->launches += 2;
    }
    CU_OK(cudaMemsetAsync(sl.flag.p, 0, 16, sl.stream));
    if (det && det->harris) {
      HarrisWorkspace hw = plan.hw;
      hw.det = ws;
      hw.scores = sl.h_scores.as<int>(); hw.pts = sl.h_pts.as<HPoint>(); hw.keep = sl.h_keep.as<uint8_t>();
      hw.sorted = sl.h_sorted.as<HPoint>(); hw.layer_kept = sl.h_layer_kept.as<int>(); hw.occ = sl.h_occ.as<uint8_t>();
      hw.surv = sl.h_surv.as<HPoint>(); hw.layer_surv = sl.h_layer_surv.as<int>();
      tm.mark(3); tm.mark(4);
      if (det->harris == 2) {
        CU_OK(launch_harris_legacy_detect(g, hw, c, det->radius, d_kps, d_counts, cap, sl.flag.as<int>(), sl.stream));
        ctx->launches += 6;
      } else {
        CU_OK(launch_harris_detect(g, hw, c, det->radius, det->abs_thr, det->max_kpt, d_kps, d_counts, cap, sl.flag.as<int>(), sl.stream));
        ctx->launches += 3 * g.n_layers + 5;
      }
    } else if (det) {
      CU_OK(launch_agast_detect(g, ws, c, det->thresh, sl.stream));
      ctx->launches += g.n_layers;
      tm.mark(3);
      CU_OK(launch_corner_lists(g, ws, c, sl.flag.as<int>(), sl.stream));
      ctx->launches += 1 + g.n_layers;
      if (det->thresh < 20) {
        CU_OK(launch_corner_score_check(g, ws, c, sl.flag.as<int>() + 1, sl.stream));  // its own flag word
        ctx->launches += 1;
      }
      tm.mark(4);
      CU_OK(launch_agast_nms(g, ws, c, d_masks, mask_fs, mask_pitch, d_kps, d_counts, cap, sl.flag.as<int>(), sl.stream));
      ctx->launches += 5;
    } else {
      tm.mark(3); tm.mark(4);
    }
    tm.mark(5);
    if (ext) {
      const uint8_t* l0 = ws.pyr + g.L[0].off;
      if (integral_aside) CU_OK(cudaStreamWaitEvent(sl.stream, sl.join, 0));
      else {
        CU_OK(launch_integral(l0, g.frame_elems, g.L[0].pitch, w, h, c, sl.integral.as<int32_t>(), sl.stream));
        ctx->launches += 2;
      }
      tm.mark(6);
      // The reference samples a tightly packed image (stride == cols) and a few of its reads land one
      // column past the row end, i.e. on the first pixel of the next row; give the sampler the same
      // layout when the pitched plane differs from it.
      const uint8_t* simg = l0;
      long long sstride = g.frame_elems;
      int spitch = g.L[0].pitch;
      if (spitch != w && staged_tight) {
        simg = sl.tight.as<uint8_t>(); sstride = (long long)w * h; spitch = w;
      } else if (spitch != w) {
        const size_t fs = (size_t)w * h + 64;
        CU_OK(launch_copy_tight(l0, g.frame_elems, g.L[0].pitch, w, h, c, sl.tight.as<uint8_t>(), (long long)fs, sl.stream));
        ctx->launches += 1;
        simg = sl.tight.as<uint8_t>(); sstride = (long long)fs; spitch = w;
      }
      CU_OK(launch_describe(ext->dev, simg, sstride, spitch, w, h, c, sl.integral.as<int32_t>(), d_kps, d_counts, cap,
                            sl.kps_scratch.as<KeyPoint>(), sl.scales.as<int>(), d_desc, sl.stream));
      ctx->launches += 2;
    } else {
      tm.mark(6);
    }
    tm.mark(7);
    if (plan.n_slots == 2) CU_OK(cudaEventRecord(sl.computed, sl.stream));
    // counts + error flag to pinned host memory; the row copies follow in finish()
    CU_OK(cudaMemcpyAsync(sl.h_counts, d_counts, (size_t)c * 4, cudaMemcpyDeviceToHost, sl.stream));
    CU_OK(cudaMemcpyAsync(sl.h_counts + plan.chunk, sl.flag.p, 8, cudaMemcpyDeviceToHost, sl.stream));
    // timing mode: the number of raw AGAST corners of every frame (last entry of its layer_start row), for the byte counts of the bench
    const bool want_corners = ctx->timing && det && !det->harris;
    if (want_corners)
      CU_OK(cudaMemcpy2DAsync(sl.h_counts + plan.chunk + 4, 4, ws.layer_start + g.n_layers, (size_t)(kMaxLayers + 1) * 4, 4, (size_t)c,
                              cudaMemcpyDeviceToHost, sl.stream));
    tm.mark(8);
    {
      brisk_ctx::Pending& pd = pend[si];
      pd.f0 = f0; pd.c = c; pd.active = true; pd.corners = want_corners; pd.d_kps = d_kps; pd.d_desc = d_desc;
      pd.kps = kps; pd.counts = counts; pd.desc = ext ? desc : nullptr; pd.cap = cap; pd.desc_bytes = desc_bytes; pd.chunk = plan.chunk;
      pd.corner_cap = plan.ws.corner_cap; pd.kps_dev = kps_dev; pd.counts_dev = counts_dev; pd.desc_dev = desc_dev; pd.st = status;
    }
    // drain the OTHER slot while this chunk computes
    if (plan.n_slots == 2 && pend[si ^ 1].active) { rc = finish(si ^ 1); if (rc) return rc; }
  }
  if (async) {
    // everything but the last chunk is collected; that one stays in flight until the next chained call or brisk_sync
    for (int si = 0; si < plan.n_slots; ++si)
      if (si != last_slot) { rc = finish(si); if (rc) return rc; }
    // everything collected so far -- the previous call's last chunk and this call's chunks but the last -- is in the
    // caller's buffers when the call returns: wait for the result copies of the other slot, and for those this slot issued
    // before its last chunk was queued behind them (the event is re-recorded by every collected chunk, so it stands for
    // the latest one; a slot that never collected anything has an unrecorded event, which counts as complete)
    for (int si = 0; si < plan.n_slots; ++si) CU_OK(cudaEventSynchronize(ctx->slots[si].done));
    ctx->deferred = pend[last_slot];
    ctx->deferred_slot = last_slot;
    ctx->deferred_key = key;
    return BRISK_OK;
  }
  for (int si = 0; si < plan.n_slots; ++si) { rc = finish(si); if (rc) return rc; }
  for (int si = 0; si < plan.n_slots; ++si) {
    CU_OK(cudaStreamSynchronize(ctx->slots[si].stream));
    collect_timing(si);
  }
  return report_status(ctx, local_status);
}

}  // namespace

extern "C" {

int brisk_ctx_create(int device, void* stream, brisk_ctx** out) {
  if (!out) return BRISK_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) { cudaGetLastError(); return BRISK_ERR_CUDA; }
  if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return BRISK_ERR_CUDA; }
  brisk_ctx* ctx = new brisk_ctx;
  ctx->device = device;
  if (stream) { ctx->stream = reinterpret_cast<cudaStream_t>(stream); ctx->own_stream = false; }
  else {
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); delete ctx; return BRISK_ERR_CUDA; }
    ctx->own_stream = true;
  }
  for (auto& e : ctx->ev) cudaEventCreate(&e);
  cudaEventCreateWithFlags(&ctx->entry, cudaEventDisableTiming);
  for (Slot& sl : ctx->slots) {
    if (cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); brisk_ctx_destroy(ctx); return BRISK_ERR_CUDA; }
    if (cudaStreamCreateWithFlags(&sl.side, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); brisk_ctx_destroy(ctx); return BRISK_ERR_CUDA; }
    cudaEventCreateWithFlags(&sl.fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&sl.join, cudaEventDisableTiming);
    for (auto& e : sl.ev) cudaEventCreate(&e);
    cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&sl.computed, cudaEventDisableTiming);
  }
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
    cudaGetLastError();
    brisk_ctx_destroy(ctx);
    return BRISK_ERR_CUDA;
  }
  ctx->encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  *out = ctx;
  return BRISK_OK;
}

void brisk_ctx_destroy(brisk_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->deferred.active) drain_async(ctx);   // an asynchronous call's last chunk still goes to the caller's buffers
  cudaStreamSynchronize(ctx->stream);
  for (Slot& sl : ctx->slots) {
    if (sl.stream) cudaStreamSynchronize(sl.stream);
    for (DevBuf* b : sl.all) b->release();
    for (auto& e : sl.ev) if (e) cudaEventDestroy(e);
    if (sl.done) cudaEventDestroy(sl.done);
    if (sl.computed) cudaEventDestroy(sl.computed);
    if (sl.h_counts) cudaFreeHost(sl.h_counts);
    if (sl.stream) cudaStreamDestroy(sl.stream);
    if (sl.side) cudaStreamDestroy(sl.side);
    if (sl.fork) cudaEventDestroy(sl.fork);
    if (sl.join) cudaEventDestroy(sl.join);
  }
  DevBuf* bufs[] = {&ctx->knn_q, &ctx->knn_t, &ctx->knn_qx, &ctx->knn_tx, &ctx->knn_keys, &ctx->knn_part, &ctx->knn_idx, &ctx->knn_dist,
                    &ctx->knn_mask, &ctx->rad_counts, &ctx->rad_offsets, &ctx->rad_matches};
  for (DevBuf* b : bufs) b->release();
  if (ctx->entry) cudaEventDestroy(ctx->entry);
  for (auto& e : ctx->ev) if (e) cudaEventDestroy(e);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* brisk_last_error(const brisk_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int brisk_sync(brisk_ctx* ctx) {
  if (!ctx) return BRISK_ERR_INVALID;
  const int rc = drain_async(ctx);
  CU_OK(cudaStreamSynchronize(ctx->stream));
  return rc;
}

int brisk_ctx_set_workspace_limit(brisk_ctx* ctx, size_t bytes) {
  if (!ctx || bytes < ((size_t)64 << 20)) return BRISK_ERR_INVALID;
  ctx->ws_limit = bytes;
  return BRISK_OK;
}

int brisk_ctx_set_knn_variant(brisk_ctx* ctx, int variant) {
  if (!ctx || variant < 0 || variant > 3) return BRISK_ERR_INVALID;
  ctx->knn_variant = variant;
  return BRISK_OK;
}

int brisk_ctx_set_pipelining(brisk_ctx* ctx, int enable) {
  if (!ctx) return BRISK_ERR_INVALID;
  ctx->pipelining = enable != 0;
  return BRISK_OK;
}

int brisk_ctx_enable_timing(brisk_ctx* ctx, int enable) {
  if (!ctx) return BRISK_ERR_INVALID;
  ctx->timing = enable != 0;
  return BRISK_OK;
}

int brisk_ctx_last_timing(brisk_ctx* ctx, float* ms, int64_t* launches) {
  if (!ctx) return BRISK_ERR_INVALID;
  if (ms) memcpy(ms, ctx->ms, sizeof(ctx->ms));
  if (launches) *launches = ctx->launches;
  return BRISK_OK;
}

int brisk_ctx_last_raw_corners(brisk_ctx* ctx, int64_t* raw_corners) {
  if (!ctx || !raw_corners) return BRISK_ERR_INVALID;
  *raw_corners = ctx->raw_corners;
  return BRISK_OK;
}

int brisk_agast_detector_create(brisk_ctx* ctx, int thresh, int octaves, int suppress, brisk_detector** out) {
  if (!ctx || !out) return BRISK_ERR_INVALID;
  *out = new brisk_detector{ctx, thresh, octaves, suppress, 0};
  return BRISK_OK;
}

int brisk_harris_detector_create(brisk_ctx* ctx, int octaves, double uniformity_radius, double absolute_threshold,
                                 int64_t max_kpts, brisk_detector** out) {
  if (!ctx || !out) return BRISK_ERR_INVALID;
  brisk_detector* d = new brisk_detector{ctx, 0, octaves, 1, 0};
  d->harris = 1; d->radius = uniformity_radius; d->abs_thr = absolute_threshold; d->max_kpt = max_kpts < 0 ? -1 : (long long)max_kpts;
  *out = d;
  return BRISK_OK;
}

int brisk_harris_legacy_detector_create(brisk_ctx* ctx, double radius, brisk_detector** out) {
  if (!ctx || !out) return BRISK_ERR_INVALID;
  brisk_detector* d = new brisk_detector{ctx, 0, 0, 1, 0};
  d->harris = 2; d->radius = radius; d->abs_thr = 64; d->max_kpt = -1;
  *out = d;
  return BRISK_OK;
}

void brisk_detector_destroy(brisk_detector* det) { delete det; }

int brisk_detector_set_corner_capacity(brisk_detector* det, int corners_per_frame) {
  if (!det || corners_per_frame < 0) return BRISK_ERR_INVALID;
  det->corner_cap = corners_per_frame;
  return BRISK_OK;
}

int brisk_extractor_create(brisk_ctx* ctx, int rot, int scale, int version, float pattern_scale, const char* pattern_file,
                           brisk_extractor** out) {
  if (!ctx || !out) return BRISK_ERR_INVALID;
  *out = nullptr;
  CU_OK(cudaSetDevice(ctx->device));
  brisk_extractor* ext = new brisk_extractor;
  ext->ctx = ctx;
  ext->device = ctx->device;
  const std::string msg = build_pattern(version, pattern_scale, pattern_file, &ext->host);
  if (!msg.empty()) { delete ext; return fail(ctx, BRISK_ERR_INVALID, msg); }
  const PatternHost& ph = ext->host;
  if (ph.n_points > 96 || ph.desc_bytes <= 0 || ph.desc_bytes > 256) { delete ext; return fail(ctx, BRISK_ERR_UNSUPPORTED, "pattern too large"); }
  // Device layouts (the host tables keep the reference's): points as (x, y) pairs, sigma -- which does not depend on the
  // rotation -- next to the two normalisation constants of its point, long pairs packed into eight bytes.
  const size_t P = (size_t)ph.n_points, n_long = ph.long_pairs.size() / 4;
  std::vector<float> xy((size_t)64 * 1024 * P * 2);
  std::vector<int> consts((size_t)64 * P * 4, 0);
  for (size_t sc = 0; sc < 64; ++sc)
    for (size_t th = 0; th < 1024; ++th)
      for (size_t i = 0; i < P; ++i) {
        const float* src = &ph.points[((sc * 1024 + th) * P + i) * 3];
        float* dst = &xy[((sc * 1024 + th) * P + i) * 2];
        dst[0] = src[0]; dst[1] = src[1];
        if (memcmp(&src[2], &ph.points[(sc * 1024 * P + i) * 3 + 2], 4) != 0) { delete ext; return fail(ctx, BRISK_ERR_CUDA, "internal error: pattern sigma depends on the rotation"); }
      }
  for (size_t sc = 0; sc < 64; ++sc)
    for (size_t i = 0; i < P; ++i) {
      int* c = &consts[(sc * P + i) * 4];
      c[0] = ph.sample_consts[(sc * P + i) * 2]; c[1] = ph.sample_consts[(sc * P + i) * 2 + 1];
      memcpy(&c[2], &ph.points[(sc * 1024 * P + i) * 3 + 2], 4);
    }
  std::vector<int> lpairs(n_long * 2);
  for (size_t k = 0; k < n_long; ++k) {
    const int* lp = &ph.long_pairs[4 * k];
    if (lp[2] < -32768 || lp[2] > 32767 || lp[3] < -32768 || lp[3] > 32767) { delete ext; return fail(ctx, BRISK_ERR_UNSUPPORTED, "pattern with long pairs closer than 1/16 pixel"); }
    lpairs[2 * k] = lp[0] | (lp[1] << 16);
    lpairs[2 * k + 1] = (lp[2] & 0xffff) | (int)((unsigned)lp[3] << 16);
  }
  struct Up { DevBuf* b; const void* src; size_t bytes; };
  const Up ups[] = {{&ext->points, xy.data(), xy.size() * 4}, {&ext->size_list, ph.size_list, sizeof(ph.size_list)},
                    {&ext->short_pairs, ph.short_pairs.data(), ph.short_pairs.size() * 2}, {&ext->long_pairs, lpairs.data(), lpairs.size() * 4},
                    {&ext->breaks, ph.scale_breaks, sizeof(ph.scale_breaks)}, {&ext->consts, consts.data(), consts.size() * 4}};
  for (const Up& u : ups) {
    cudaError_t e = u.b->ensure(std::max<size_t>(u.bytes, 16));
    if (e == cudaSuccess && u.bytes) e = cudaMemcpy(u.b->p, u.src, u.bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaGetLastError(); brisk_extractor_destroy(ext); return fail(ctx, BRISK_ERR_CUDA, cudaGetErrorString(e)); }
  }
  PatternDev& d = ext->dev;
  d.points = ext->points.as<float2>(); d.size_list = ext->size_list.as<unsigned int>();
  d.short_pairs = ext->short_pairs.as<unsigned short>(); d.long_pairs = ext->long_pairs.as<int2>();
  d.scale_breaks = ext->breaks.as<float>();
  d.sample_consts = ext->consts.as<int4>();
  d.n_points = ph.n_points; d.n_short = (int)ph.short_pairs.size() / 2; d.n_long = (int)ph.long_pairs.size() / 4;
  d.desc_bytes = ph.desc_bytes; d.rot_inv = rot != 0; d.scale_inv = scale != 0; d.basic_scale = ph.basic_scale;
  *out = ext;
  return BRISK_OK;
}

void brisk_extractor_destroy(brisk_extractor* ext) {
  if (!ext) return;
  cudaSetDevice(ext->device);
  ext->points.release(); ext->size_list.release(); ext->short_pairs.release(); ext->long_pairs.release(); ext->breaks.release(); ext->consts.release();
  delete ext;
}

int brisk_extractor_descriptor_size(const brisk_extractor* ext) { return ext ? ext->dev.desc_bytes : BRISK_ERR_INVALID; }

int brisk_extractor_pattern(const brisk_extractor* ext, int32_t counts[4], float* points_xys, float* scale_list,
                            uint32_t* size_list, uint32_t* short_pairs, int32_t* long_pairs) {
  if (!ext) return BRISK_ERR_INVALID;
  const PatternHost& ph = ext->host;
  if (counts) { counts[0] = ph.n_points; counts[1] = (int)ph.short_pairs.size() / 2; counts[2] = (int)ph.long_pairs.size() / 4; counts[3] = ph.desc_bytes; }
  if (points_xys) memcpy(points_xys, ph.points.data(), ph.points.size() * 4);
  if (scale_list) memcpy(scale_list, ph.scale_list, sizeof(ph.scale_list));
  if (size_list) memcpy(size_list, ph.size_list, sizeof(ph.size_list));
  if (short_pairs) for (size_t i = 0; i < ph.short_pairs.size(); ++i) short_pairs[i] = ph.short_pairs[i];
  if (long_pairs) memcpy(long_pairs, ph.long_pairs.data(), ph.long_pairs.size() * 4);
  return BRISK_OK;
}

int brisk_detect(brisk_ctx* ctx, brisk_detector* det, const uint8_t* imgs, int n, int w, int h, size_t stride,
                 size_t frame_pitch, const uint8_t* masks, brisk_keypoint* kps, int32_t* counts, int cap) {
  if (!ctx || !det || det->ctx != ctx) return fail(ctx, BRISK_ERR_INVALID, "detector does not belong to this context");
  return run_batch(ctx, det, nullptr, imgs, n, w, h, stride, frame_pitch, masks, kps, counts, cap, nullptr);
}

int brisk_describe(brisk_ctx* ctx, brisk_extractor* ext, const uint8_t* imgs, int n, int w, int h, size_t stride,
                   size_t frame_pitch, brisk_keypoint* kps, int32_t* counts, int cap, uint8_t* desc) {
  if (!ctx || !ext || ext->ctx != ctx) return fail(ctx, BRISK_ERR_INVALID, "extractor does not belong to this context");
  return run_batch(ctx, nullptr, ext, imgs, n, w, h, stride, frame_pitch, nullptr, kps, counts, cap, desc);
}

int brisk_detect_describe(brisk_ctx* ctx, brisk_detector* det, brisk_extractor* ext, const uint8_t* imgs, int n, int w,
                          int h, size_t stride, size_t frame_pitch, const uint8_t* masks, brisk_keypoint* kps,
                          int32_t* counts, int cap, uint8_t* desc) {
  if (!ctx || !det || !ext || det->ctx != ctx || ext->ctx != ctx) return fail(ctx, BRISK_ERR_INVALID, "detector / extractor do not belong to this context");
  return run_batch(ctx, det, ext, imgs, n, w, h, stride, frame_pitch, masks, kps, counts, cap, desc);
}

int brisk_detect_describe_async(brisk_ctx* ctx, brisk_detector* det, brisk_extractor* ext, const uint8_t* imgs, int n, int w,
                                int h, size_t stride, size_t frame_pitch, const uint8_t* masks, brisk_keypoint* kps,
                                int32_t* counts, int cap, uint8_t* desc) {
  if (!ctx || !det || !ext || det->ctx != ctx || ext->ctx != ctx) return fail(ctx, BRISK_ERR_INVALID, "detector / extractor do not belong to this context");
  return run_batch(ctx, det, ext, imgs, n, w, h, stride, frame_pitch, masks, kps, counts, cap, desc, true);
}

int brisk_compute_scale(brisk_ctx* ctx, brisk_detector* det, const uint8_t* imgs, int n, int w, int h, size_t stride,
                        size_t frame_pitch, const brisk_keypoint* kps_in, const int32_t* counts_in, int cap_in,
                        brisk_keypoint* kps_out, int32_t* counts_out, int cap_out) {
  int rc = check_image_args(ctx, imgs, n, w, h, stride, frame_pitch);
  if (rc) return rc;
  if (!det || det->harris) return fail(ctx, BRISK_ERR_INVALID, "ComputeScale needs a BriskFeatureDetector (AGAST) handle");
  if (!kps_in || !counts_in || cap_in <= 0 || !kps_out || !counts_out || cap_out <= 0) return fail(ctx, BRISK_ERR_INVALID, "bad key point arguments");
  if (det->octaves < 0 || 2 * det->octaves > kMaxLayers) return fail(ctx, BRISK_ERR_UNSUPPORTED, "octaves must be in [0, 6]");
  if (!det->suppress && det->octaves != 0)
    return fail(ctx, BRISK_ERR_UNSUPPORTED, "suppressScaleNonmaxima=false is only defined for octaves == 0 (the reference reads out of bounds otherwise)");
  CU_OK(cudaSetDevice(ctx->device));
  memset(ctx->ms, 0, sizeof(ctx->ms));
  ctx->launches = 0;
  if (n == 0) return BRISK_OK;
  {
    PyramidGeom probe;
    build_geom(w, h, det->octaves, &probe);
    for (int i = 0; i < probe.n_layers; ++i)
      if (probe.L[i].w < 8 || probe.L[i].h < 8) return fail(ctx, BRISK_ERR_INVALID, "image too small: every pyramid layer must be at least 8x8");
  }
  // the largest input list sizes the per-frame scratch (one slot per layer and provided point); an empty vector makes
  // the reference detect instead (with the lower threshold of the map set to 0)
  KpLists io{kps_in, counts_in, cap_in, kps_out, counts_out, cap_out};
  rc = io.init(ctx, n, "ComputeScale without key points runs the detector in the reference; use detect()");
  if (rc) return rc;
  const int in_max = io.in_max;
  brisk_detector sized = *det;
  const int n_layers = det->octaves == 0 ? 1 : 2 * det->octaves;
  if ((long long)n_layers * in_max > (1ll << 26)) return fail(ctx, BRISK_ERR_UNSUPPORTED, "too many provided key points");
  // one slot per layer and provided point, plus room for the corners of layers that keep no point (there the
  // reference runs its detector): the detector's own corner capacity
  int detect_cap = det->corner_cap;
  if (detect_cap <= 0) detect_cap = std::min(std::max((int)(((long long)w * h) / 8), 4096), 1 << 20);  // no lower bound: more corners
  sized.corner_cap = n_layers * in_max + detect_cap;
  Plan plan;
  rc = make_plan(ctx, &sized, nullptr, n, w, h, cap_out, &plan, false, false);
  if (rc) return rc;
  const PyramidGeom& g = plan.g;
  Slot& sl = ctx->slots[0];
  const size_t c_max = (size_t)plan.chunk;
  rc = io.reserve(ctx, sl, c_max);
  if (rc) return rc;
  if (!is_device_ptr(imgs)) CU_OK(sl.tight.ensure(c_max * ((size_t)w * h + 64)));
  CU_OK(sl.prov_base.ensure(c_max * (kMaxLayers + 1) * 4));
  std::vector<int> h_kept(c_max * kTieStride);
  CU_OK(cudaEventRecord(ctx->entry, ctx->stream));
  CU_OK(cudaStreamWaitEvent(sl.stream, ctx->entry, 0));
  bool empty_layer = false, corner_overflow = false;
  const DetectWorkspace ws = slot_ws(plan, sl);
  for (int f0 = 0; f0 < n; f0 += plan.chunk) {
    const int c = std::min(plan.chunk, n - f0);
    rc = io.upload(ctx, sl, f0, c);
    if (rc) return rc;
    CUtensorMap map;
    int write_l0 = 0;
    bool staged_tight = false;
    rc = stage_input(ctx, sl, plan, imgs + (size_t)f0 * frame_pitch, c, w, h, stride, frame_pitch, &map, &write_l0, &staged_tight);
    if (rc) return rc;
    CU_OK(launch_pyramid(map, g, ws.pyr, c, write_l0, sl.stream));
    CU_OK(cudaMemsetAsync(sl.flag.p, 0, 16, sl.stream));
    // which layers keep none of their frame's points?  Those run the detector (threshold map without lower bound).
    CU_OK(launch_provided_count(g, ws, c, io.d_in, io.d_cin, cap_in, in_max, sl.stream));
    CU_OK(cudaMemcpyAsync(h_kept.data(), ws.n_ties, (size_t)c * kTieStride * 4, cudaMemcpyDeviceToHost, sl.stream));
    CU_OK(cudaStreamSynchronize(sl.stream));
    int with_fallback = 0;
    for (int f = 0; f < c; ++f)
      for (int l = 0; l < g.n_layers; ++l) with_fallback |= h_kept[(size_t)f * kTieStride + l] == 0;
    if (with_fallback) {
      CU_OK(launch_agast_detect(g, ws, c, det->thresh, sl.stream, 0));
      CU_OK(launch_corner_lists(g, ws, c, sl.flag.as<int>(), sl.stream));
      ctx->launches += 2 * g.n_layers + 3;
    }
    CU_OK(launch_provided_scale(g, ws, c, io.d_in, io.d_cin, cap_in, in_max, with_fallback, sl.prov_base.as<int>(), io.d_out, io.d_cout,
                                cap_out, sl.flag.as<int>(), sl.stream));
    ctx->launches += 7;
    int flag = 0;
    rc = io.download(ctx, sl, plan.chunk, f0, c, &flag);
    if (rc) return rc;
    if (flag == 3) empty_layer = true;
    else if (flag) corner_overflow = true;
  }
  if (empty_layer) return fail(ctx, BRISK_ERR_CUDA, "internal error: a layer without provided key points was not detected on");
  if (corner_overflow) return fail(ctx, BRISK_ERR_CAPACITY, "raw corner capacity exceeded on a layer without provided key points; raise it with brisk_detector_set_corner_capacity");
  if (io.truncated) return fail(ctx, BRISK_ERR_CAPACITY, "key point capacity (cap_out) exceeded; counts hold the true numbers");
  return BRISK_OK;
}

int brisk_harris_detect_passed(brisk_ctx* ctx, brisk_detector* det, int n, int w, int h, const brisk_keypoint* kps_in,
                               const int32_t* counts_in, int cap_in, brisk_keypoint* kps_out, int32_t* counts_out, int cap_out) {
  if (!ctx) return BRISK_ERR_INVALID;
  if (!det || !det->harris) return fail(ctx, BRISK_ERR_INVALID, "passed key points are a mode of the Harris scale-space detector");
  if (n < 0 || w <= 0 || h <= 0 || w > 8191 || h > 8191) return fail(ctx, BRISK_ERR_INVALID, "bad image size");
  if (!kps_in || !counts_in || cap_in <= 0 || !kps_out || !counts_out || cap_out <= 0) return fail(ctx, BRISK_ERR_INVALID, "bad key point arguments");
  // a second layer would index its (smaller) occupancy map with layer 0's image coordinates: out of bounds in the reference
  if (det->octaves != 0) return fail(ctx, BRISK_ERR_UNSUPPORTED, "passed key points are only defined for octaves == 0 (the reference writes out of bounds otherwise)");
  if (!(det->radius > 0.0)) {
    if (det->max_kpt <= 0) return fail(ctx, BRISK_ERR_UNSUPPORTED, "key point bucketing (uniformityRadius <= 0) needs a finite maxNumKpt > 0");
  } else if (!occupancy_fits(det->radius, w, h)) return fail(ctx, BRISK_ERR_UNSUPPORTED, kOccupancyTooLarge);
  if (w < 8 || h < 8) return fail(ctx, BRISK_ERR_INVALID, "image too small");
  CU_OK(cudaSetDevice(ctx->device));
  memset(ctx->ms, 0, sizeof(ctx->ms));
  ctx->launches = 0;
  if (n == 0) return BRISK_OK;
  KpLists io{kps_in, counts_in, cap_in, kps_out, counts_out, cap_out};
  int rc = io.init(ctx, n, "an empty key point vector means detection; use detect()");
  if (rc) return rc;
  const int in_max = io.in_max;
  brisk_detector sized = *det;
  sized.corner_cap = in_max;
  Plan plan;
  rc = make_plan(ctx, &sized, nullptr, n, w, h, cap_out, &plan, false, false);
  if (rc) return rc;
  const PyramidGeom& g = plan.g;
  Slot& sl = ctx->slots[0];
  rc = io.reserve(ctx, sl, (size_t)plan.chunk);
  if (rc) return rc;
  CU_OK(cudaEventRecord(ctx->entry, ctx->stream));
  CU_OK(cudaStreamWaitEvent(sl.stream, ctx->entry, 0));
  HarrisWorkspace hw = plan.hw;
  hw.det = slot_ws(plan, sl);
  hw.scores = sl.h_scores.as<int>(); hw.pts = sl.h_pts.as<HPoint>(); hw.keep = sl.h_keep.as<uint8_t>();
  hw.sorted = sl.h_sorted.as<HPoint>(); hw.layer_kept = sl.h_layer_kept.as<int>(); hw.occ = sl.h_occ.as<uint8_t>();
  hw.surv = sl.h_surv.as<HPoint>(); hw.layer_surv = sl.h_layer_surv.as<int>();
  bool bad_point = false;
  for (int f0 = 0; f0 < n; f0 += plan.chunk) {
    const int c = std::min(plan.chunk, n - f0);
    rc = io.upload(ctx, sl, f0, c);
    if (rc) return rc;
    CU_OK(cudaMemsetAsync(sl.flag.p, 0, 16, sl.stream));
    CU_OK(launch_harris_passed(g, hw, c, det->radius, det->max_kpt, io.d_in, io.d_cin, cap_in, in_max, io.d_out, io.d_cout, cap_out,
                               sl.flag.as<int>(), sl.stream));
    ctx->launches += 4;
    int flag = 0;
    rc = io.download(ctx, sl, plan.chunk, f0, c, &flag);
    if (rc) return rc;
    if (flag == 4) bad_point = true;
  }
  if (bad_point) return fail(ctx, BRISK_ERR_INVALID, "a passed key point with response > 1e6 lies outside the image (or its response does not fit an int)");
  if (io.truncated) return fail(ctx, BRISK_ERR_CAPACITY, "key point capacity (cap_out) exceeded; counts hold the true numbers");
  return BRISK_OK;
}

int brisk_harris_scores(brisk_ctx* ctx, const uint8_t* img, int w, int h, size_t stride, int abs_threshold, int32_t* scores,
                        int32_t* maxima_sxy, int cap, int32_t* n_maxima) {
  int rc = check_image_args(ctx, img, 1, w, h, stride, stride * (size_t)h);
  if (rc) return rc;
  if (w < 8 || h < 8) return fail(ctx, BRISK_ERR_INVALID, "image too small");
  if ((maxima_sxy && (cap <= 0 || !n_maxima)) || (scores && is_device_ptr(scores)) || (maxima_sxy && is_device_ptr(maxima_sxy)))
    return fail(ctx, BRISK_ERR_INVALID, "bad output arguments (host buffers expected)");
  CU_OK(cudaSetDevice(ctx->device));
  brisk_detector det{ctx, 0, 0, 1, 0};
  det.harris = 1; det.radius = 0.0; det.abs_thr = abs_threshold; det.max_kpt = 1;
  det.corner_cap = maxima_sxy ? cap : 0;
  Plan plan;
  rc = make_plan(ctx, &det, nullptr, 1, w, h, 1, &plan, false, false);
  if (rc) return rc;
  const PyramidGeom& g = plan.g;
  Slot& sl = ctx->slots[0];
  if (!is_device_ptr(img)) CU_OK(sl.tight.ensure((size_t)w * h + 64));
  CU_OK(cudaEventRecord(ctx->entry, ctx->stream));
  CU_OK(cudaStreamWaitEvent(sl.stream, ctx->entry, 0));
  HarrisWorkspace hw = plan.hw;
  hw.det = slot_ws(plan, sl);
  hw.scores = sl.h_scores.as<int>(); hw.pts = sl.h_pts.as<HPoint>();
  CUtensorMap map; int write_l0 = 0; bool staged = false;
  rc = stage_input(ctx, sl, plan, img, 1, w, h, stride, stride * (size_t)h, &map, &write_l0, &staged);
  if (rc) return rc;
  CU_OK(launch_pyramid(map, g, hw.det.pyr, 1, write_l0, sl.stream));
  CU_OK(cudaMemsetAsync(sl.flag.p, 0, 16, sl.stream));
  CU_OK(launch_harris_score_maxima(g, hw, 1, abs_threshold, sl.flag.as<int>(), sl.stream));
  ctx->launches = 5;
  if (scores) CU_OK(cudaMemcpy2DAsync(scores, (size_t)w * 4, hw.scores + g.L[0].off, (size_t)g.L[0].pitch * 4, (size_t)w * 4, h, cudaMemcpyDeviceToHost, sl.stream));
  int ls[2] = {0, 0};   // one layer: its first slot and the total
  CU_OK(cudaMemcpyAsync(ls, hw.det.layer_start, sizeof(ls), cudaMemcpyDeviceToHost, sl.stream));
  CU_OK(cudaStreamSynchronize(sl.stream));
  const int total = ls[1];
  if (n_maxima) *n_maxima = total;
  if (maxima_sxy) {
    const int m = std::min(total, cap);
    std::vector<HPoint> pts((size_t)std::max(m, 1));
    if (m > 0) CU_OK(cudaMemcpy(pts.data(), hw.pts, (size_t)m * sizeof(HPoint), cudaMemcpyDeviceToHost));
    for (int i = 0; i < m; ++i) { maxima_sxy[3 * i] = pts[i].score; maxima_sxy[3 * i + 1] = pts[i].x; maxima_sxy[3 * i + 2] = pts[i].y; }
    if (total > cap) return fail(ctx, BRISK_ERR_CAPACITY, "more 2-D maxima than the output capacity; n_maxima holds the number needed");
  }
  return BRISK_OK;
}

// Halfsample16 / Twothirdsample16 on host or device buffers (strides in BYTES, like cv::Mat::step).
static int sample16(brisk_ctx* ctx, int two_third, const uint16_t* src, int w, int h, size_t src_stride, uint16_t* dst, size_t dst_stride) {
  if (!ctx) return BRISK_ERR_INVALID;
  const int dw = two_third ? 2 * (w / 3) : w / 2, dh = two_third ? 2 * (h / 3) : h / 2;
  if (!src || !dst || w <= 0 || h <= 0 || src_stride < (size_t)w * 2 || dst_stride < (size_t)dw * 2 || src_stride % 2 || dst_stride % 2)
    return fail(ctx, BRISK_ERR_INVALID, "bad 16-bit image arguments");
  // narrower images never enter the reference's SSE loops: it writes nothing (image-down-sampling.cc:70-75, 407-411)
  if (two_third ? (w / 3) * 3 < 12 : w < 16) return fail(ctx, BRISK_ERR_UNSUPPORTED, two_third ? "Twothirdsample16 needs at least 12 columns" : "Halfsample16 needs at least 16 columns");
  if (dw == 0 || dh == 0) return BRISK_OK;
  CU_OK(cudaSetDevice(ctx->device));
  const bool sdev = is_device_ptr(src), ddev = is_device_ptr(dst);
  const uint16_t* ds = src; uint16_t* dd = dst;
  long long sp = (long long)src_stride / 2, dp = (long long)dst_stride / 2;
  if (!sdev) {
    CU_OK(ctx->knn_q.ensure((size_t)w * h * 2));
    CU_OK(cudaMemcpy2DAsync(ctx->knn_q.p, (size_t)w * 2, src, src_stride, (size_t)w * 2, h, cudaMemcpyHostToDevice, ctx->stream));
    ds = ctx->knn_q.as<uint16_t>(); sp = w;
  }
  if (!ddev) {
    CU_OK(ctx->knn_t.ensure((size_t)dw * dh * 2));
    dd = ctx->knn_t.as<uint16_t>(); dp = dw;
  }
  if (two_third) CU_OK(launch_twothirdsample16(ds, w, h, sp, dd, dp, ctx->stream));
  else CU_OK(launch_halfsample16(ds, w, h, sp, dd, dp, ctx->stream));
  ctx->launches = 1;
  if (!ddev) CU_OK(cudaMemcpy2DAsync(dst, dst_stride, dd, (size_t)dw * 2, (size_t)dw * 2, dh, cudaMemcpyDeviceToHost, ctx->stream));
  CU_OK(cudaStreamSynchronize(ctx->stream));
  return BRISK_OK;
}

int brisk_halfsample16(brisk_ctx* ctx, const uint16_t* src, int w, int h, size_t src_stride, uint16_t* dst, size_t dst_stride) {
  return sample16(ctx, 0, src, w, h, src_stride, dst, dst_stride);
}

int brisk_twothirdsample16(brisk_ctx* ctx, const uint16_t* src, int w, int h, size_t src_stride, uint16_t* dst, size_t dst_stride) {
  return sample16(ctx, 1, src, w, h, src_stride, dst, dst_stride);
}

int brisk_debug_pyramid(brisk_ctx* ctx, int octaves, const uint8_t* img, int w, int h, size_t stride, uint8_t* out,
                        int32_t* dims, int* n_layers) {
  int rc = check_image_args(ctx, img, 1, w, h, stride, stride * (size_t)h);
  if (rc) return rc;
  if (octaves < 0 || 2 * octaves > kMaxLayers) return fail(ctx, BRISK_ERR_UNSUPPORTED, "octaves must be in [0, 6]");
  CU_OK(cudaSetDevice(ctx->device));
  brisk_detector det{ctx, 60, octaves, 1, 4096};
  Plan plan;
  rc = make_plan(ctx, &det, nullptr, 1, w, h, 1, &plan);
  if (rc) return rc;
  Slot& sl = ctx->slots[0];
  const DetectWorkspace ws = slot_ws(plan, sl);
  (void)ws;
  CUtensorMap map; int write_l0;
  rc = stage_input(ctx, sl, plan, img, 1, w, h, stride, stride * (size_t)h, &map, &write_l0);
  if (rc) return rc;
  CU_OK(launch_pyramid(map, plan.g, ws.pyr, 1, write_l0, sl.stream));
  size_t off = 0;
  for (int i = 0; i < plan.g.n_layers; ++i) {
    const LayerGeom& L = plan.g.L[i];
    if (dims) { dims[2 * i] = L.w; dims[2 * i + 1] = L.h; }
    if (out && L.w > 0 && L.h > 0)
      CU_OK(cudaMemcpy2DAsync(out + off, L.w, ws.pyr + L.off, L.pitch, L.w, L.h, cudaMemcpyDeviceToHost, sl.stream));
    off += (size_t)L.w * L.h;
  }
  if (n_layers) *n_layers = plan.g.n_layers;
  CU_OK(cudaStreamSynchronize(sl.stream));
  return BRISK_OK;
}

int brisk_debug_integral(brisk_ctx* ctx, const uint8_t* img, int w, int h, size_t stride, int32_t* out) {
  int rc = check_image_args(ctx, img, 1, w, h, stride, stride * (size_t)h);
  if (rc) return rc;
  if (!out) return fail(ctx, BRISK_ERR_INVALID, "null output");
  CU_OK(cudaSetDevice(ctx->device));
  if (is_device_ptr(img) || is_device_ptr(out)) return fail(ctx, BRISK_ERR_INVALID, "the integral-image check takes host buffers");
  Plan plan;
  rc = make_plan(ctx, nullptr, nullptr, 1, w, h, 1, &plan);
  if (rc) return rc;
  Slot& sl = ctx->slots[0];
  const DetectWorkspace ws = slot_ws(plan, sl);
  (void)ws;
  CU_OK(sl.integral.ensure(((size_t)w * h * 4 + (size_t)integral_aux_elems(w, h)) * 4));
  CUtensorMap map; int write_l0;
  rc = stage_input(ctx, sl, plan, img, 1, w, h, stride, stride * (size_t)h, &map, &write_l0);
  if (rc) return rc;
  CU_OK(launch_pyramid(map, plan.g, ws.pyr, 1, write_l0, sl.stream));
  CU_OK(launch_integral(ws.pyr + plan.g.L[0].off, plan.g.frame_elems, plan.g.L[0].pitch, w, h, 1, sl.integral.as<int32_t>(), sl.stream));
  // the device holds one encoded block per pixel (describe_logic.cuh): rebuild S from the d components and check
  // that the other fields -- a, b, c, the block's own pixel and the pixel one row up / one column right in the tightly
  // packed image -- say the same
  std::vector<int32_t> blk((size_t)w * h * 4);
  CU_OK(cudaMemcpyAsync(blk.data(), sl.integral.p, blk.size() * 4, cudaMemcpyDeviceToHost, sl.stream));
  CU_OK(cudaStreamSynchronize(sl.stream));
  const size_t iw = (size_t)w + 1;
  for (size_t i = 0; i < iw * ((size_t)h + 1); ++i) out[i] = 0;
  for (int Y = 0; Y < h; ++Y)
    for (int X = 0; X < w; ++X) out[(size_t)(Y + 1) * iw + X + 1] = blk[((size_t)Y * w + X) * 4 + 3];
  auto pixel = [&](long long i) { return (i >= 0 && i < (long long)w * h) ? (int)img[(size_t)(i / w) * stride + (size_t)(i % w)] : 0; };
  for (int Y = 0; Y < h; ++Y)
    for (int X = 0; X < w; ++X) {
      const int32_t* e = &blk[((size_t)Y * w + X) * 4];
      const BlockFields b = decode_block(Block4{e[0], e[1], e[2], e[3]});
      const int up_right = Y ? pixel((long long)(Y - 1) * w + X + 1) : 0;
      if (b.a != out[(size_t)Y * iw + X] || b.b != out[(size_t)Y * iw + X + 1] || b.c != out[(size_t)(Y + 1) * iw + X] ||
          b.pix != pixel((long long)Y * w + X) || b.up_right != up_right)
        return fail(ctx, BRISK_ERR_CUDA, "internal error: inconsistent integral-image blocks");
    }
  return BRISK_OK;
}

int brisk_debug_corners(brisk_ctx* ctx, brisk_detector* det, const uint8_t* img, int w, int h, size_t stride,
                        int32_t* corners_xys, int cap, int32_t* layer_counts) {
  int rc = check_image_args(ctx, img, 1, w, h, stride, stride * (size_t)h);
  if (rc) return rc;
  if (!det || !corners_xys) return fail(ctx, BRISK_ERR_INVALID, "null argument");
  CU_OK(cudaSetDevice(ctx->device));
  Plan plan;
  rc = make_plan(ctx, det, nullptr, 1, w, h, 1, &plan);
  if (rc) return rc;
  Slot& sl = ctx->slots[0];
  const DetectWorkspace ws = slot_ws(plan, sl);
  (void)ws;
  CUtensorMap map; int write_l0;
  rc = stage_input(ctx, sl, plan, img, 1, w, h, stride, stride * (size_t)h, &map, &write_l0);
  if (rc) return rc;
  CU_OK(launch_pyramid(map, plan.g, ws.pyr, 1, write_l0, sl.stream));
  CU_OK(cudaMemsetAsync(sl.flag.p, 0, 16, sl.stream));
  CU_OK(launch_agast_detect(plan.g, ws, 1, det->thresh, sl.stream));
  CU_OK(launch_corner_lists(plan.g, ws, 1, sl.flag.as<int>(), sl.stream));
  std::vector<int> ls(kMaxLayers + 1);
  CU_OK(cudaMemcpyAsync(ls.data(), ws.layer_start, ls.size() * 4, cudaMemcpyDeviceToHost, sl.stream));
  CU_OK(cudaStreamSynchronize(sl.stream));
  const int total = std::min(ls[plan.g.n_layers], ws.corner_cap);
  std::vector<uint32_t> packed(std::max(total, 1));
  std::vector<uint16_t> cm((size_t)plan.g.frame_elems);
  CU_OK(cudaMemcpy(packed.data(), ws.corners, (size_t)total * 4, cudaMemcpyDeviceToHost));
  CU_OK(cudaMemcpy(cm.data(), ws.cm, cm.size() * 2, cudaMemcpyDeviceToHost));
  for (int i = 0; i < total && i < cap; ++i) {
    const int x = packed[i] & 0x1fff, y = (packed[i] >> 13) & 0x1fff, l = packed[i] >> 26;
    corners_xys[3 * i] = x; corners_xys[3 * i + 1] = y;
    corners_xys[3 * i + 2] = cm[(size_t)plan.g.L[l].off + (size_t)y * plan.g.L[l].pitch + x] & kCmT;
  }
  if (layer_counts) for (int l = 0; l < plan.g.n_layers; ++l) layer_counts[l] = ls[l + 1] - ls[l];
  return total;
}

int brisk_debug_scores(brisk_ctx* ctx, const uint8_t* img, int w, int h, size_t stride, uint8_t* out916, uint8_t* out58) {
  int rc = check_image_args(ctx, img, 1, w, h, stride, stride * (size_t)h);
  if (rc) return rc;
  if (!out916 || !out58) return fail(ctx, BRISK_ERR_INVALID, "null output");
  CU_OK(cudaSetDevice(ctx->device));
  Plan plan;
  rc = make_plan(ctx, nullptr, nullptr, 1, w, h, 1, &plan);
  if (rc) return rc;
  Slot& sl = ctx->slots[0];
  const DetectWorkspace ws = slot_ws(plan, sl);
  (void)ws;
  CU_OK(sl.kps.ensure((size_t)w * h * 2));
  CUtensorMap map; int write_l0;
  rc = stage_input(ctx, sl, plan, img, 1, w, h, stride, stride * (size_t)h, &map, &write_l0);
  if (rc) return rc;
  CU_OK(launch_pyramid(map, plan.g, ws.pyr, 1, write_l0, sl.stream));
  uint8_t* d916 = sl.kps.as<uint8_t>();
  uint8_t* d58 = d916 + (size_t)w * h;
  CU_OK(launch_dense_scores(plan.g.L[0], ws.pyr + plan.g.L[0].off, d916, d58, sl.stream));
  CU_OK(cudaMemcpyAsync(out916, d916, (size_t)w * h, cudaMemcpyDeviceToHost, sl.stream));
  CU_OK(cudaMemcpyAsync(out58, d58, (size_t)w * h, cudaMemcpyDeviceToHost, sl.stream));
  CU_OK(cudaStreamSynchronize(sl.stream));
  return BRISK_OK;
}

int brisk_debug_nms_state(brisk_ctx* ctx, brisk_detector* det, const uint8_t* img, int w, int h, size_t stride,
                          uint16_t* cm_out, uint8_t* bm_out) {
  if (!ctx || !det) return BRISK_ERR_INVALID;
  std::vector<brisk_keypoint> kps(1 << 16);
  int32_t count = 0;
  int rc = run_batch(ctx, det, nullptr, img, 1, w, h, stride, stride * (size_t)h, nullptr, kps.data(), &count, (int)kps.size(), nullptr);
  if (rc) return rc;
  PyramidGeom g;
  build_geom(w, h, det->octaves, &g);
  size_t off = 0;
  for (int i = 0; i < g.n_layers; ++i) {
    const LayerGeom& L = g.L[i];
    if (cm_out) CU_OK(cudaMemcpy2D(cm_out + off, (size_t)L.w * 2, ctx->slots[0].cm.as<uint16_t>() + L.off, (size_t)L.pitch * 2, (size_t)L.w * 2, L.h, cudaMemcpyDeviceToHost));
    if (bm_out) CU_OK(cudaMemcpy2D(bm_out + off, L.w, ctx->slots[0].bm.as<uint8_t>() + L.off, L.pitch, L.w, L.h, cudaMemcpyDeviceToHost));
    off += (size_t)L.w * L.h;
  }
  return count;
}

int brisk_debug_nms_ties(brisk_ctx* ctx, int32_t* ties /* [12] of frame 0 of the last detect call */) {
  if (!ctx || !ties || !ctx->slots[0].rounds.p) return BRISK_ERR_INVALID;
  CU_OK(cudaMemcpy(ties, ctx->slots[0].rounds.p, kMaxLayers * 4, cudaMemcpyDeviceToHost));
  return BRISK_OK;
}

// ---------------------------------------------------------------------------
// Hamming matching.
// ---------------------------------------------------------------------------

// Tensor map over expanded descriptor rows ([rows][kbytes] signed bytes): boxes of 128 rows x 128 bytes, 128B swizzle.
static int encode_rows_map(brisk_ctx* ctx, const void* base, long long rows, int kbytes, int box_rows, CUtensorMap* map) {
  cuuint64_t dims[2] = {(cuuint64_t)kbytes, (cuuint64_t)std::max<long long>(rows, box_rows)};
  cuuint64_t strides[1] = {(cuuint64_t)kbytes};
  cuuint32_t box[2] = {128, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = ctx->encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ctx, BRISK_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
  return BRISK_OK;
}

static int knn_keys_impl(brisk_ctx* ctx, const uint8_t* query, int64_t nq, const uint8_t* train, int64_t nt, int desc_bytes,
                         int k, int64_t offset, unsigned long long** keys_out, int* kr_out) {
  if (!ctx) return BRISK_ERR_INVALID;
  if (!query || !train || nq < 0 || nt < 0 || k < 1) return fail(ctx, BRISK_ERR_INVALID, "bad kNN arguments");
  if (desc_bytes < 4 || desc_bytes % 4 || desc_bytes > 496) return fail(ctx, BRISK_ERR_UNSUPPORTED, "descriptor rows must be a multiple of 4 bytes, at most 496");
  if (offset + nt > 0xffffffffll) return fail(ctx, BRISK_ERR_UNSUPPORTED, "train index does not fit 32 bits");
  if (k > 65536) return fail(ctx, BRISK_ERR_UNSUPPORTED, "k too large");
  // the extractors' row widths and k <= 8 have their own kernels; anything else takes the general one
  const bool general = (desc_bytes != 48 && desc_bytes != 64 && desc_bytes != 128) || k > 8;
  CU_OK(cudaSetDevice(ctx->device));
  const uint8_t* dq = query; const uint8_t* dt = train;
  if (!is_device_ptr(query) || (uintptr_t)query % 16) {
    CU_OK(ctx->knn_q.ensure(std::max<size_t>((size_t)nq * desc_bytes, 16)));
    CU_OK(cudaMemcpyAsync(ctx->knn_q.p, query, (size_t)nq * desc_bytes, cudaMemcpyDefault, ctx->stream));
    dq = ctx->knn_q.as<uint8_t>();
  }
  if (!is_device_ptr(train) || (uintptr_t)train % 16) {
    CU_OK(ctx->knn_t.ensure(std::max<size_t>((size_t)nt * desc_bytes, 16)));
    CU_OK(cudaMemcpyAsync(ctx->knn_t.p, train, (size_t)nt * desc_bytes, cudaMemcpyDefault, ctx->stream));
    dt = ctx->knn_t.as<uint8_t>();
  }
  if (general) {
    const int kr = knn_any_round_k(k);
    CU_OK(ctx->knn_keys.ensure(std::max<size_t>((size_t)nq * kr * 8, 16)));
    if (ctx->timing) cudaEventRecord(ctx->ev[0], ctx->stream);
    CU_OK(launch_hamming_knn_any(dq, nq, dt, nt, desc_bytes, k, offset, nullptr, ctx->knn_keys.as<unsigned long long>(), ctx->stream));
    if (ctx->timing) cudaEventRecord(ctx->ev[1], ctx->stream);
    ctx->launches = kr / 8;
    *keys_out = ctx->knn_keys.as<unsigned long long>();
    *kr_out = kr;
    return BRISK_OK;
  }
  const int kr = knn_round_k(k);
  const bool tensor = k == 2 && (desc_bytes == 48 || desc_bytes == 64);
  const bool mx4 = ctx->knn_variant == 3 && tensor && nq > 0 && nt > 0;   // FP4 form
  const bool mma = ctx->knn_variant == 1 && tensor, tc5 = (ctx->knn_variant == 2 || (ctx->knn_variant == 3 && !mx4)) && tensor && nq > 0 && nt > 0;
  const int splits = mx4 ? knn_tc5_num_splits(nq, nt, knn_tc5mx_query_tiles()) : tc5 ? knn_tc5_num_splits(nq, nt) : (mma ? knn_mma_num_splits(nq, nt) : knn_num_splits(nq, nt));
  CU_OK(ctx->knn_keys.ensure(std::max<size_t>((size_t)nq * kr * 8, 16)));
  if (splits > 1) CU_OK(ctx->knn_part.ensure((size_t)splits * nq * kr * 8));
  if (ctx->timing) cudaEventRecord(ctx->ev[0], ctx->stream);
  if (mx4) {
    // descriptor bits -> E2M1 +-1.0 once per call (4x the rows in HBM), then the block-scaled FP4 tcgen05 kernel
    CU_OK(ctx->knn_tx.ensure(knn_tc5mx_expanded_bytes(nt, desc_bytes)));
    CU_OK(launch_expand_e2m1(dt, nt, desc_bytes, ctx->knn_tx.as<uint8_t>(), ctx->stream));
    CU_OK(ctx->knn_qx.ensure(knn_tc5mx_expanded_bytes(nq, desc_bytes)));
    CU_OK(launch_expand_e2m1(dq, nq, desc_bytes, ctx->knn_qx.as<uint8_t>(), ctx->stream));
    CUtensorMap mq, mt;
    int rc = encode_rows_map(ctx, ctx->knn_tx.p, nt, desc_bytes * 4, knn_tc5mx_tile_rows(), &mt);
    if (rc) return rc;
    rc = encode_rows_map(ctx, ctx->knn_qx.p, nq, desc_bytes * 4, 128, &mq);
    if (rc) return rc;
    CU_OK(launch_hamming_knn2_tc5mx(mq, nq, mt, nt, desc_bytes, offset, ctx->knn_keys.as<unsigned long long>(),
                                    ctx->knn_part.as<unsigned long long>(), splits, ctx->stream));
    ctx->launches = 2;
  } else if (tc5) {
    // descriptor bits -> signed bytes once per call (8x the rows in HBM), then the tcgen05 kernel on TMA-staged tiles
    const bool ts = knn_tc5_queries_in_tmem() != 0;
    CU_OK(ctx->knn_tx.ensure(knn_tc5_expanded_bytes(nt, desc_bytes)));
    CU_OK(launch_expand_pm1(dt, nt, desc_bytes, ctx->knn_tx.as<uint8_t>(), ctx->stream));
    CUtensorMap mq, mt;
    int rc = encode_rows_map(ctx, ctx->knn_tx.p, nt, desc_bytes * 8, ts ? 128 : knn_tc5_tile_rows(), &mt);
    if (rc) return rc;
    if (ts) {
      // queries stay packed: the kernel's epilogue threads expand their own row into tensor memory
      CU_OK(launch_hamming_knn2_tc5ts(dq, nq, mt, nt, desc_bytes, offset, ctx->knn_keys.as<unsigned long long>(),
                                      ctx->knn_part.as<unsigned long long>(), splits, ctx->stream));
      ctx->launches = 1;
    } else {
      CU_OK(ctx->knn_qx.ensure(knn_tc5_expanded_bytes(nq, desc_bytes)));
      CU_OK(launch_expand_pm1(dq, nq, desc_bytes, ctx->knn_qx.as<uint8_t>(), ctx->stream));
      rc = encode_rows_map(ctx, ctx->knn_qx.p, nq, desc_bytes * 8, 128, &mq);
      if (rc) return rc;
      CU_OK(launch_hamming_knn2_tc5(mq, nq, mt, nt, desc_bytes, offset, ctx->knn_keys.as<unsigned long long>(),
                                    ctx->knn_part.as<unsigned long long>(), splits, ctx->stream));
      ctx->launches = 2;
    }
  } else if (mma)
    CU_OK(launch_hamming_knn2_mma(dq, nq, dt, nt, desc_bytes, offset, ctx->knn_keys.as<unsigned long long>(),
                                  ctx->knn_part.as<unsigned long long>(), splits, ctx->stream));
  else
    CU_OK(launch_hamming_knn_ex(dq, nq, dt, nt, desc_bytes, k, offset, ctx->knn_keys.as<unsigned long long>(),
                                ctx->knn_part.as<unsigned long long>(), splits, ctx->stream));
  if (ctx->timing) cudaEventRecord(ctx->ev[1], ctx->stream);
  ctx->launches += splits > 1 ? 2 : 1;
  *keys_out = ctx->knn_keys.as<unsigned long long>();
  *kr_out = kr;
  return BRISK_OK;
}

// keys [nq][kr] (device) -> idx/dist [nq][k] in host or device memory
static int knn_emit(brisk_ctx* ctx, const unsigned long long* keys, int64_t nq, int kr, int k, int32_t* idx, int32_t* dist) {
  if (!idx || !dist) return fail(ctx, BRISK_ERR_INVALID, "null output");
  if (nq == 0) return BRISK_OK;
  CU_OK(ctx->knn_idx.ensure((size_t)nq * kr * 4));
  CU_OK(ctx->knn_dist.ensure((size_t)nq * kr * 4));
  CU_OK(launch_knn_unpack(keys, nq * kr, ctx->knn_idx.as<int32_t>(), ctx->knn_dist.as<int32_t>(), ctx->stream));
  ctx->launches += 1;
  CU_OK(cudaMemcpy2DAsync(idx, (size_t)k * 4, ctx->knn_idx.p, (size_t)kr * 4, (size_t)k * 4, (size_t)nq, cudaMemcpyDefault, ctx->stream));
  CU_OK(cudaMemcpy2DAsync(dist, (size_t)k * 4, ctx->knn_dist.p, (size_t)kr * 4, (size_t)k * 4, (size_t)nq, cudaMemcpyDefault, ctx->stream));
  CU_OK(cudaStreamSynchronize(ctx->stream));
  if (ctx->timing) {
    float ms = 0;
    memset(ctx->ms, 0, sizeof(ctx->ms));
    if (cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]) == cudaSuccess) ctx->ms[BRISK_STAGE_KNN] = ms; else cudaGetLastError();
  }
  return BRISK_OK;
}

int brisk_hamming_knn(brisk_ctx* ctx, const uint8_t* query, int64_t nq, const uint8_t* train, int64_t nt, int desc_bytes,
                      int k, int32_t* idx, int32_t* dist) {
  unsigned long long* keys = nullptr; int kr = 0;
  int rc = knn_keys_impl(ctx, query, nq, train, nt, desc_bytes, k, 0, &keys, &kr);
  if (rc) return rc;
  return knn_emit(ctx, keys, nq, kr, k, idx, dist);
}

// query / train / mask staged on the device (16-byte aligned rows) for the masked and radius paths
static int stage_match_inputs(brisk_ctx* ctx, const uint8_t* query, int64_t nq, const uint8_t* train, int64_t nt, int desc_bytes,
                              const uint8_t* mask, const uint8_t** dq, const uint8_t** dt, const uint8_t** dm) {
  if (!query || !train || nq < 0 || nt < 0) return fail(ctx, BRISK_ERR_INVALID, "bad matcher arguments");
  if (desc_bytes < 4 || desc_bytes % 4 || desc_bytes > 496) return fail(ctx, BRISK_ERR_UNSUPPORTED, "descriptor rows must be a multiple of 4 bytes, at most 496");
  if (nt > 0x7fffffffll) return fail(ctx, BRISK_ERR_UNSUPPORTED, "train index does not fit 31 bits");
  CU_OK(cudaSetDevice(ctx->device));
  *dq = query; *dt = train; *dm = mask;
  if (!is_device_ptr(query) || (uintptr_t)query % 16) {
    CU_OK(ctx->knn_q.ensure(std::max<size_t>((size_t)nq * desc_bytes, 16)));
    CU_OK(cudaMemcpyAsync(ctx->knn_q.p, query, (size_t)nq * desc_bytes, cudaMemcpyDefault, ctx->stream));
    *dq = ctx->knn_q.as<uint8_t>();
  }
  if (!is_device_ptr(train) || (uintptr_t)train % 16) {
    CU_OK(ctx->knn_t.ensure(std::max<size_t>((size_t)nt * desc_bytes, 16)));
    CU_OK(cudaMemcpyAsync(ctx->knn_t.p, train, (size_t)nt * desc_bytes, cudaMemcpyDefault, ctx->stream));
    *dt = ctx->knn_t.as<uint8_t>();
  }
  if (mask && !is_device_ptr(mask)) {
    CU_OK(ctx->knn_mask.ensure(std::max<size_t>((size_t)nq * nt, 16)));
    CU_OK(cudaMemcpyAsync(ctx->knn_mask.p, mask, (size_t)nq * nt, cudaMemcpyDefault, ctx->stream));
    *dm = ctx->knn_mask.as<uint8_t>();
  }
  return BRISK_OK;
}

int brisk_hamming_knn_masked(brisk_ctx* ctx, const uint8_t* query, int64_t nq, const uint8_t* train, int64_t nt, int desc_bytes,
                             int k, const uint8_t* mask, int32_t* idx, int32_t* dist) {
  if (!ctx) return BRISK_ERR_INVALID;
  if (!mask) return brisk_hamming_knn(ctx, query, nq, train, nt, desc_bytes, k, idx, dist);
  if (k < 1 || k > 65536) return fail(ctx, BRISK_ERR_INVALID, "bad kNN arguments");
  const uint8_t *dq, *dt, *dm;
  int rc = stage_match_inputs(ctx, query, nq, train, nt, desc_bytes, mask, &dq, &dt, &dm);
  if (rc) return rc;
  const bool general = (desc_bytes != 48 && desc_bytes != 64 && desc_bytes != 128) || k > 8;
  const int kr = general ? knn_any_round_k(k) : knn_round_k(k);
  CU_OK(ctx->knn_keys.ensure(std::max<size_t>((size_t)nq * kr * 8, 16)));
  if (ctx->timing) cudaEventRecord(ctx->ev[0], ctx->stream);
  if (general) CU_OK(launch_hamming_knn_any(dq, nq, dt, nt, desc_bytes, k, 0, dm, ctx->knn_keys.as<unsigned long long>(), ctx->stream));
  else CU_OK(launch_hamming_knn_masked(dq, nq, dt, nt, desc_bytes, k, dm, ctx->knn_keys.as<unsigned long long>(), ctx->stream));
  if (ctx->timing) cudaEventRecord(ctx->ev[1], ctx->stream);
  ctx->launches = 1;
  return knn_emit(ctx, ctx->knn_keys.as<unsigned long long>(), nq, kr, k, idx, dist);
}

int brisk_hamming_radius(brisk_ctx* ctx, const uint8_t* query, int64_t nq, const uint8_t* train, int64_t nt, int desc_bytes,
                         float max_distance, const uint8_t* mask, int sort, int64_t* offsets, int32_t* idx, int32_t* dist,
                         int64_t capacity) {
  if (!ctx) return BRISK_ERR_INVALID;
  if (!offsets || capacity < 0 || (capacity > 0 && (!idx || !dist))) return fail(ctx, BRISK_ERR_INVALID, "bad radius-match arguments");
  const uint8_t *dq, *dt, *dm;
  int rc = stage_match_inputs(ctx, query, nq, train, nt, desc_bytes, mask, &dq, &dt, &dm);
  if (rc) return rc;
  CU_OK(ctx->rad_counts.ensure(std::max<size_t>((size_t)nq * 8, 16)));
  CU_OK(ctx->rad_offsets.ensure((size_t)(nq + 1) * 8));
  long long* d_off = ctx->rad_offsets.as<long long>();
  if (nq == 0) CU_OK(cudaMemsetAsync(d_off, 0, 8, ctx->stream));
  if (ctx->timing) cudaEventRecord(ctx->ev[0], ctx->stream);
  CU_OK(launch_hamming_radius(0, dq, nq, dt, nt, desc_bytes, max_distance, dm, ctx->rad_counts.as<long long>(), d_off, nullptr, 0, 0, ctx->stream));
  long long total = 0;
  CU_OK(cudaMemcpyAsync(&total, d_off + nq, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CU_OK(cudaStreamSynchronize(ctx->stream));
  ctx->launches = 2;
  const long long emit = std::min<long long>(total, capacity);
  if (emit > 0) {
    CU_OK(ctx->rad_matches.ensure((size_t)emit * 8));
    CU_OK(launch_hamming_radius(1, dq, nq, dt, nt, desc_bytes, max_distance, dm, nullptr, d_off, ctx->rad_matches.p, emit,
                                total <= capacity ? sort : 0, ctx->stream));
    CU_OK(ctx->knn_idx.ensure((size_t)emit * 4));
    CU_OK(ctx->knn_dist.ensure((size_t)emit * 4));
    CU_OK(launch_radius_unpack(ctx->rad_matches.p, emit, ctx->knn_idx.as<int32_t>(), ctx->knn_dist.as<int32_t>(), ctx->stream));
    ctx->launches += sort ? 3 : 2;
    CU_OK(cudaMemcpyAsync(idx, ctx->knn_idx.p, (size_t)emit * 4, cudaMemcpyDefault, ctx->stream));
    CU_OK(cudaMemcpyAsync(dist, ctx->knn_dist.p, (size_t)emit * 4, cudaMemcpyDefault, ctx->stream));
  }
  if (ctx->timing) cudaEventRecord(ctx->ev[1], ctx->stream);
  CU_OK(cudaMemcpyAsync(offsets, d_off, (size_t)(nq + 1) * 8, cudaMemcpyDefault, ctx->stream));
  CU_OK(cudaStreamSynchronize(ctx->stream));
  if (ctx->timing) {
    float ms = 0;
    memset(ctx->ms, 0, sizeof(ctx->ms));
    if (cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]) == cudaSuccess) ctx->ms[BRISK_STAGE_KNN] = ms; else cudaGetLastError();
  }
  if (total > capacity) return fail(ctx, BRISK_ERR_CAPACITY, "more radius matches than the output capacity; offsets[nq] holds the number needed");
  return BRISK_OK;
}

// std::sort of one query's DMatch list by distance (brute-force-matcher.cc:160,210), as libstdc++ leaves it: lists of
// more than 16 entries go through introsort, which permutes entries of equal distance.  Host only.
int brisk_std_sort_matches(int64_t n, int32_t* train_idx, int32_t* img_idx, float* distance) {
  if (n < 0 || (n > 0 && (!train_idx || !img_idx || !distance)) || n > 0x7fffffff) return BRISK_ERR_INVALID;
  struct M { float d; int32_t t, i; };
  struct Less { bool operator()(const M& a, const M& b) const { return a.d < b.d; } };
  std::vector<M> v((size_t)n);
  for (int64_t j = 0; j < n; ++j) v[(size_t)j] = M{distance[j], train_idx[j], img_idx[j]};
  gs_sort(Less(), v.data(), (int)n);
  for (int64_t j = 0; j < n; ++j) { distance[j] = v[(size_t)j].d; train_idx[j] = v[(size_t)j].t; img_idx[j] = v[(size_t)j].i; }
  return BRISK_OK;
}

int brisk_hamming_knn_keys(brisk_ctx* ctx, const uint8_t* query, int64_t nq, const uint8_t* train_shard, int64_t nt,
                           int desc_bytes, int k, int64_t global_train_offset, uint64_t* keys_dev) {
  if (!keys_dev || !is_device_ptr(keys_dev)) return fail(ctx, BRISK_ERR_INVALID, "keys_dev must be device memory");
  if (k != knn_round_k(k) || (desc_bytes != 48 && desc_bytes != 64 && desc_bytes != 128))
    return fail(ctx, BRISK_ERR_UNSUPPORTED, "sharded kNN supports k in {1, 2, 4, 8} and rows of 48, 64 or 128 bytes");
  unsigned long long* keys = nullptr; int kr = 0;
  int rc = knn_keys_impl(ctx, query, nq, train_shard, nt, desc_bytes, k, global_train_offset, &keys, &kr);
  if (rc) return rc;
  CU_OK(cudaMemcpyAsync(keys_dev, keys, (size_t)nq * kr * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  CU_OK(cudaStreamSynchronize(ctx->stream));
  return BRISK_OK;
}

int brisk_knn_merge_keys(brisk_ctx* ctx, const uint64_t* gathered_keys_dev, int n_shards, int64_t nq, int k, int32_t* idx,
                         int32_t* dist) {
  if (!ctx) return BRISK_ERR_INVALID;
  if (!gathered_keys_dev || n_shards < 1 || nq < 0 || k != knn_round_k(k) || k > 8) return fail(ctx, BRISK_ERR_INVALID, "bad merge arguments");
  CU_OK(cudaSetDevice(ctx->device));
  CU_OK(ctx->knn_keys.ensure(std::max<size_t>((size_t)nq * k * 8, 16)));
  CU_OK(launch_knn_merge(reinterpret_cast<const unsigned long long*>(gathered_keys_dev), n_shards, nq, k, ctx->knn_keys.as<unsigned long long>(), ctx->stream));
  ctx->launches = 1;
  if (ctx->timing) { cudaEventRecord(ctx->ev[0], ctx->stream); cudaEventRecord(ctx->ev[1], ctx->stream); }
  return knn_emit(ctx, ctx->knn_keys.as<unsigned long long>(), nq, k, k, idx, dist);
}

int brisk_hamming_distance(brisk_ctx* ctx, const uint8_t* a, const uint8_t* b, int64_t n, int desc_bytes, int32_t* dist) {
  if (!ctx || !a || !b || !dist || n < 0 || desc_bytes <= 0) return fail(ctx, BRISK_ERR_INVALID, "bad arguments");
  CU_OK(cudaSetDevice(ctx->device));
  if (n == 0) return BRISK_OK;
  const size_t bytes = (size_t)n * desc_bytes;
  const uint8_t *da = a, *db = b;
  if (!is_device_ptr(a)) {
    CU_OK(ctx->knn_q.ensure(std::max<size_t>(bytes, 16)));
    CU_OK(cudaMemcpyAsync(ctx->knn_q.p, a, bytes, cudaMemcpyHostToDevice, ctx->stream));
    da = ctx->knn_q.as<uint8_t>();
  }
  if (!is_device_ptr(b)) {
    CU_OK(ctx->knn_t.ensure(std::max<size_t>(bytes, 16)));
    CU_OK(cudaMemcpyAsync(ctx->knn_t.p, b, bytes, cudaMemcpyHostToDevice, ctx->stream));
    db = ctx->knn_t.as<uint8_t>();
  }
  const bool out_dev = is_device_ptr(dist);
  int32_t* dd = dist;
  if (!out_dev) {
    CU_OK(ctx->knn_dist.ensure((size_t)n * 4));
    dd = ctx->knn_dist.as<int32_t>();
  }
  CU_OK(launch_hamming_pairs(da, db, n, desc_bytes, dd, ctx->stream));
  if (!out_dev) CU_OK(cudaMemcpyAsync(dist, dd, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU_OK(cudaStreamSynchronize(ctx->stream));
  return BRISK_OK;
}

}  // extern "C"
