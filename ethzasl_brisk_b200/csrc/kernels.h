// Host-side launchers of the sm_100a kernels (internal to libbrisk_b200.so).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "brisk_common.cuh"

namespace briskb200 {

// Per-batch device workspace of the AGAST detector; plane blocks are laid out
// [frame][layer] following PyramidGeom.
constexpr int kTieStride = 16;
struct DetectWorkspace {
  uint8_t* pyr;         // u8 image planes
  uint16_t* cm;         // u16 corner maps
  uint8_t* bm;          // u8 touch maps
  int* rowcnt;          // [frame][total_rows] corners per row -> exclusive prefix (in place)
  int* layer_start;     // [frame][kMaxLayers + 1] first corner slot of each layer
  uint32_t* corners;    // [frame][corner_cap] packed x | y << 13 | layer << 26, layer-major raster order
  uint8_t* fwin;        // [frame][corner_cap][32] 5x5 FAST scores of tying corners
  float* checks;        // [frame][corner_cap][8]  CheckResult (32 bytes)
  KeyPoint* kp_tmp;     // [frame][corner_cap]
  uint8_t* kp_valid;    // [frame][corner_cap]
  int* n_ties;          // [frame][kTieStride]: tying corners per layer, then [kMaxLayers] = corners that need the scale checks
  int* surv;            // [frame][corner_cap] slots of the corners that passed IsMax2D's comparisons (any order)
  int total_rows;       // sum of layer heights
  int row_off[kMaxLayers + 1];
  int corner_cap;       // per frame
};

cudaError_t launch_pyramid(const CUtensorMap& src_map, const PyramidGeom& g, uint8_t* pyr, int n_frames, int write_l0,
                           cudaStream_t stream);

// 16-bit samplers (pitches in elements); the caller checks the reference's size limits (w >= 16 resp. 3 * (w / 3) >= 12).
cudaError_t launch_halfsample16(const uint16_t* src, int w, int h, long long spitch, uint16_t* dst, long long dpitch, cudaStream_t stream);
cudaError_t launch_twothirdsample16(const uint16_t* src, int w, int h, long long spitch, uint16_t* dst, long long dpitch, cudaStream_t stream);

// Threshold map + AGAST 9-16 segment test -> corner map + per-row counts.
// `lower`: lower bound of the threshold map, kLowerThreshold (detection) or 0 (ComputeScale's pyramid).
cudaError_t launch_agast_detect(const PyramidGeom& g, const DetectWorkspace& ws, int n_frames, int thresh, cudaStream_t stream,
                                int lower = kLowerThreshold);
// Row prefix sums and ordered corner lists.
cudaError_t launch_corner_lists(const PyramidGeom& g, const DetectWorkspace& ws, int n_frames, int* overflow_flag,
                                cudaStream_t stream);
// After the corner lists: *flag = 5 when a detected corner scores <= 2 (possible only for thresh < 20; such a corner's
// cache entry would not be sticky and the NMS closed form would not apply).
cudaError_t launch_corner_score_check(const PyramidGeom& g, const DetectWorkspace& ws, int n_frames, int* flag, cudaStream_t stream);
// Scale-space NMS + refinement -> ordered key points [frame][kp_cap], counts[frame].
cudaError_t launch_agast_nms(const PyramidGeom& g, const DetectWorkspace& ws, int n_frames, const uint8_t* masks,
                             long long mask_frame_stride, int mask_pitch, KeyPoint* out, int* counts, int kp_cap,
                             int* error_flag, cudaStream_t stream);

// BriskFeatureDetector::ComputeScale: scale-space checks + refinement of caller-provided key points
// (in: [frame][in_cap], in_counts[frame] <= in_max).  launch_provided_count leaves in ws.n_ties[frame][layer]
// how many points each layer keeps; a layer that keeps none is a fallback layer (the reference detects there):
// if any exists the caller runs launch_agast_detect(lower = 0) + launch_corner_lists and passes with_fallback = 1
// (otherwise *error_flag is set to 3).  base: [frame][kMaxLayers + 1] scratch.  Needs ws.corner_cap >=
// n_layers * in_max (+ the corners of the fallback layers); *error_flag = 1 when that does not hold.
cudaError_t launch_provided_count(const PyramidGeom& g, const DetectWorkspace& ws, int n_frames, const KeyPoint* in,
                                  const int* in_counts, int in_cap, int in_max, cudaStream_t stream);
cudaError_t launch_provided_scale(const PyramidGeom& g, const DetectWorkspace& ws, int n_frames, const KeyPoint* in,
                                  const int* in_counts, int in_cap, int in_max, int with_fallback, int* base, KeyPoint* out,
                                  int* counts, int kp_cap, int* error_flag, cudaStream_t stream);

// Harris scale-space detector (harris.cu).
struct HPoint;
struct HarrisWorkspace {
  DetectWorkspace det;   // pyr, rowcnt, layer_start, total_rows, row_off, corner_cap (= maxima capacity per frame)
  int* scores;           // i32 score planes, PyramidGeom layout
  HPoint* pts;           // [frame][cap] 2-D maxima, layer-major raster order
  uint8_t* keep;         // [frame][cap] 3-D NMS verdicts
  HPoint* sorted;        // [frame][cap] kept maxima per layer, std::sort order
  int* layer_kept;       // [frame][kMaxLayers]
  uint8_t* occ;          // [frame][occ_frame_bytes] occupancy maps
  HPoint* surv;          // [frame][cap] survivors of the uniformity enforcement
  int* layer_surv;       // [frame][kMaxLayers]
  long long occ_frame_bytes;
  long long occ_off[kMaxLayers];
  int occ_w[kMaxLayers], occ_h[kMaxLayers];
};
cudaError_t launch_row_scan(const PyramidGeom& g, const DetectWorkspace& ws, int n_frames, int* overflow_flag, cudaStream_t stream);
cudaError_t launch_harris_score_maxima(const PyramidGeom& g, const HarrisWorkspace& hw, int n_frames, int thr, int* overflow_flag,
                                       cudaStream_t stream);
cudaError_t launch_harris_detect(const PyramidGeom& g, const HarrisWorkspace& hw, int n_frames, double radius, double abs_thr,
                                 long long max_kpt, KeyPoint* out, int* counts, int kp_cap, int* overflow_flag, cudaStream_t stream);

// Legacy single-scale HarrisFeatureDetector(radius): one layer; *error_flag = 6 when a point's occupancy indices leave the map
// (the reference then accesses memory out of bounds: nothing to match), 1 when the maxima exceed corner_cap.
cudaError_t launch_harris_legacy_detect(const PyramidGeom& g, const HarrisWorkspace& hw, int n_frames, double radius, KeyPoint* out,
                                        int* counts, int kp_cap, int* error_flag, cudaStream_t stream);

// detect() of the Harris scale-space detector on a non-empty key-point vector ("use passed key points"), one layer:
// response > 1e6 filter, std::sort replay, uniformity enforcement / bucketing, unrefined output.  *error_flag = 4 when a
// passed point lies outside the image (or its response does not fit an int).
cudaError_t launch_harris_passed(const PyramidGeom& g, const HarrisWorkspace& hw, int n_frames, double radius, long long max_kpt,
                                 const KeyPoint* in, const int* in_counts, int in_cap, int in_max, KeyPoint* out, int* counts,
                                 int kp_cap, int* error_flag, cudaStream_t stream);

cudaError_t launch_dense_scores(const LayerGeom& L, const uint8_t* img, uint8_t* out916, uint8_t* out58, cudaStream_t stream);

// Descriptor extraction.
struct PatternDev {
  const float2* points;  // [64][1024][P] x, y
  const unsigned int* size_list;  // [64]
  const unsigned short* short_pairs;  // [n_short][2] (i, j)
  const int2* long_pairs;  // [n_long] (i | j << 16, wdx & 0xffff | wdy << 16)
  const float* scale_breaks;  // [64] smallest keypoint size mapping to scale index s (s >= 1)
  const int4* sample_consts;  // [64][P] (scaling, scaling2, bits of sigma / 2, 0) of every pattern point
  int n_points, n_short, n_long, desc_bytes;
  int rot_inv, scale_inv, basic_scale;
};

// `integral` holds n_frames block images (w x h blocks of four int32: {S(Y,X), S(Y,X+1), S(Y+1,X), S(Y+1,X+1)}, S = the
// (h+1) x (w+1) integral image of the reference) followed by n_frames * integral_aux_elems(w, h) scratch ints.
long long integral_aux_elems(int w, int h);
cudaError_t launch_integral(const uint8_t* imgs, long long frame_stride, int pitch, int w, int h, int n_frames,
                            int32_t* integral, cudaStream_t stream);
// Border cull (stable, per frame) + descriptor computation.  kps is in/out
// [frame][kp_cap], counts in/out, desc out [frame][kp_cap][desc_bytes].
cudaError_t launch_copy_tight(const uint8_t* src, long long src_frame_stride, int src_pitch, int w, int h, int n_frames,
                              uint8_t* dst, long long dst_frame_stride, cudaStream_t stream);
cudaError_t launch_describe(const PatternDev& pat, const uint8_t* imgs, long long frame_stride, int pitch, int w, int h,
                            int n_frames, const int32_t* integral, KeyPoint* kps, int* counts, int kp_cap,
                            KeyPoint* kps_scratch, int* scale_scratch, uint8_t* desc, cudaStream_t stream);

// Brute-force Hamming k-nearest neighbours (k <= 8); packed (dist << 32 | train idx) u64 keys,
// sorted ascending per query, kr = knn_round_k(k) keys per query.
int knn_num_splits(long long nq, long long nt);
int knn_round_k(int k);
// Any row width (a multiple of 4 bytes up to 496; whole 128-bit words are compared) and any k: keys [nq][knn_any_round_k(k)],
// optional [nq][nt] byte mask.
int knn_any_round_k(int k);
cudaError_t launch_hamming_knn_any(const uint8_t* q, long long nq, const uint8_t* t, long long nt, int desc_bytes, int k,
                                   long long train_index_offset, const uint8_t* mask, unsigned long long* keys, cudaStream_t stream);
cudaError_t launch_hamming_knn_ex(const uint8_t* q, long long nq, const uint8_t* t, long long nt, int desc_bytes, int k,
                                  long long train_index_offset, unsigned long long* keys, unsigned long long* part,
                                  int splits, cudaStream_t stream);
// kNN with a [nq][nt] byte mask (0 = pair excluded); keys [nq][knn_round_k(k)].
cudaError_t launch_hamming_knn_masked(const uint8_t* q, long long nq, const uint8_t* t, long long nt, int desc_bytes, int k,
                                      const uint8_t* mask, unsigned long long* keys, cudaStream_t stream);
// Radius match: phase 0 -> counts[nq], offsets[nq + 1]; phase 1 -> (train index, distance) int pairs at offsets
// (train order, or the reference's std::sort order when `sort`), at most `capacity` pairs in all.
cudaError_t launch_hamming_radius(int phase, const uint8_t* q, long long nq, const uint8_t* t, long long nt, int desc_bytes,
                                  float max_distance, const uint8_t* mask, long long* counts, long long* offsets, void* matches,
                                  long long capacity, int sort, cudaStream_t stream);
cudaError_t launch_radius_unpack(const void* matches, long long n, int32_t* idx, int32_t* dist, cudaStream_t stream);
// Tensor-core (u8 IMMA on 0/1-expanded bits) variant, k == 2, 48/64-byte rows (hamming_mma.cu).
int knn_mma_num_splits(long long nq, long long nt);
cudaError_t launch_hamming_knn2_mma(const uint8_t* q, long long nq, const uint8_t* t, long long nt, int desc_bytes,
                                    long long train_index_offset, unsigned long long* keys, unsigned long long* part,
                                    int splits, cudaStream_t stream);
// tcgen05 variant (hamming_tc5.cu): kind::i8 MMAs on +-1 expanded rows, TMEM accumulators, TMA operands; k == 2, 48/64-byte rows.
// The caller expands both sets (launch_expand_pm1; knn_tc5_expanded_bytes per set) and passes tensor maps over the expanded
// rows: 2-D, dims {8 * desc_bytes, max(rows, box rows)}, box {128, 128} (queries) / {128, knn_tc5_tile_rows()} (train),
// CU_TENSOR_MAP_SWIZZLE_128B.
size_t knn_tc5_expanded_bytes(long long rows, int desc_bytes);
cudaError_t launch_expand_pm1(const uint8_t* src, long long rows, int desc_bytes, uint8_t* dst, cudaStream_t stream);
int knn_tc5_num_splits(long long nq, long long nt, int query_tiles = 0);
int knn_tc5_tile_rows();
cudaError_t launch_hamming_knn2_tc5(const CUtensorMap& map_q, long long nq, const CUtensorMap& map_t, long long nt, int desc_bytes,
                                    long long train_index_offset, unsigned long long* keys, unsigned long long* part,
                                    int splits, cudaStream_t stream);
// The same with the queries in TMEM (raw descriptor rows in, expanded by the kernel): measured alternative, BRISK_B200_TC5_MODE=ts.
int knn_tc5_queries_in_tmem();
// FP4 form (tcgen05.mma kind::mxf4.block_scale, E2M1 +-1.0 operands, unit scale factors): 64-byte rows, k = 2, train tiles
// of knn_tc5mx_tile_rows() rows; operands expanded to 4 bits per descriptor bit (256 bytes per row).
size_t knn_tc5mx_expanded_bytes(long long rows, int desc_bytes);
int knn_tc5mx_tile_rows();
int knn_tc5mx_query_tiles();
cudaError_t launch_expand_e2m1(const uint8_t* src, long long rows, int desc_bytes, uint8_t* dst, cudaStream_t stream);
cudaError_t launch_hamming_knn2_tc5mx(const CUtensorMap& map_q, long long nq, const CUtensorMap& map_t, long long nt, int desc_bytes,
                                      long long train_index_offset, unsigned long long* keys, unsigned long long* part,
                                      int splits, cudaStream_t stream);
cudaError_t launch_hamming_knn2_tc5ts(const uint8_t* q, long long nq, const CUtensorMap& map_t, long long nt, int desc_bytes,
                                      long long train_index_offset, unsigned long long* keys, unsigned long long* part,
                                      int splits, cudaStream_t stream);
// brisk::Hamming::operator() on n pairs of desc_bytes-byte rows: popcount of the XOR over desc_bytes / 16 whole 128-bit words.
cudaError_t launch_hamming_pairs(const uint8_t* a, const uint8_t* b, long long n, int desc_bytes, int32_t* dist, cudaStream_t stream);
cudaError_t launch_knn_unpack(const unsigned long long* keys, long long n, int32_t* idx, int32_t* dist, cudaStream_t stream);
cudaError_t launch_knn_merge(const unsigned long long* gathered, int n_shards, long long nq, int k, unsigned long long* out,
                             cudaStream_t stream);

}  // namespace briskb200
