/* brisk_b200 -- C ABI of the B200-native BRISK hot path (libbrisk_b200.so).
 *
 * This is the drop-in boundary: the reference (ethz-asl/ethzasl_brisk) has no
 * FFI layer, its boundary is the C++ class surface of the headers in brisk/include/brisk/.
 * The header-only C++ classes in include/brisk/ (same names, constructor
 * arguments and semantics) and the Python mirror ethzasl_brisk_b200/api.py are
 * thin hosts over these entry points.  Each entry point cites the reference
 * interface it replaces.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on
 * success or a negative brisk_status and never throws; the message of the last
 * failure is available from brisk_last_error().  Image, key-point, descriptor
 * and count buffers may live in host OR device memory (detected per pointer);
 * host buffers are copied inside the call.  A context owns one CUDA stream and
 * all workspaces: use one context per host thread / GPU.  There is no CPU
 * fallback: without a usable CUDA device every call fails.
 */
#ifndef BRISK_B200_H_
#define BRISK_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct brisk_ctx brisk_ctx;
typedef struct brisk_detector brisk_detector;
typedef struct brisk_extractor brisk_extractor;

/* Binary compatible with cv::KeyPoint (pt.x, pt.y, size, angle, response,
 * octave, class_id): 28 bytes. */
typedef struct {
  float x, y, size, angle, response;
  int32_t octave, class_id;
} brisk_keypoint;

typedef enum {
  BRISK_OK = 0,
  BRISK_ERR_INVALID = -1,     /* bad argument */
  BRISK_ERR_CUDA = -2,        /* CUDA runtime / driver failure */
  BRISK_ERR_UNSUPPORTED = -3, /* configuration outside the implemented range */
  BRISK_ERR_CAPACITY = -4     /* more raw corners / key points than the configured capacity */
} brisk_status;

/* Context: device, stream, workspaces.  `stream` is a cudaStream_t to run on
 * (e.g. the current torch stream) or NULL for a private non-blocking stream.  The default stream has handle 0 == NULL:
 * pass cudaStreamLegacy ((void*)0x1) or cudaStreamPerThread ((void*)0x2) to name it.  Work of a call is ordered after
 * what the caller queued on that stream; device inputs produced on OTHER streams must be complete before the call. */
int brisk_ctx_create(int device, void* stream, brisk_ctx** out);
void brisk_ctx_destroy(brisk_ctx* ctx);
const char* brisk_last_error(const brisk_ctx* ctx);
/* Waits for the context's work, collects what brisk_detect_describe_async left in flight and returns the status of
 * the asynchronous calls since the last brisk_sync. */
int brisk_sync(brisk_ctx* ctx);
/* Upper bound on device workspace bytes used per call (frames are processed in
 * chunks that fit); default 8 GiB. */
int brisk_ctx_set_workspace_limit(brisk_ctx* ctx, size_t bytes);
/* Matcher kernel for unmasked k == 2 searches over 48- or 64-byte rows (everything else runs XOR + POPC tiles); results are
 * identical: 0 = XOR + POPC always, 1 = mma.sync IMMA on byte-expanded bits, 2 = tcgen05.mma kind::i8 (TMEM accumulators, TMA
 * operands), 3 (default) = tcgen05.mma kind::mxf4 on +-1.0 E2M1 values. */
int brisk_ctx_set_knn_variant(brisk_ctx* ctx, int variant);
/* Chunks of a batch normally alternate between two streams so that copies and serial kernel tails
 * of one chunk overlap the kernels of the other (default on).  Off: one stream, stages back to back
 * -- used to time individual stages without overlap. */
int brisk_ctx_set_pipelining(brisk_ctx* ctx, int enable);
/* Device time in milliseconds of the kernels of the last detect / describe /
 * detect_describe / knn call, per stage (see BRISK_STAGE_*), measured with CUDA
 * events on the context stream.  Requires brisk_ctx_enable_timing(ctx, 1). */
int brisk_ctx_enable_timing(brisk_ctx* ctx, int enable);
enum { BRISK_STAGE_H2D = 0, BRISK_STAGE_PYRAMID, BRISK_STAGE_DETECT, BRISK_STAGE_LISTS, BRISK_STAGE_NMS,
       BRISK_STAGE_INTEGRAL, BRISK_STAGE_DESCRIBE, BRISK_STAGE_D2H, BRISK_STAGE_KNN, BRISK_STAGE_COUNT };
int brisk_ctx_last_timing(brisk_ctx* ctx, float* ms /* [BRISK_STAGE_COUNT] */, int64_t* launches);
/* With timing enabled: raw AGAST corners (before scale-space NMS) summed over the frames of the last detect call. */
int brisk_ctx_last_raw_corners(brisk_ctx* ctx, int64_t* raw_corners);

/* brisk::BriskFeatureDetector(int thresh, int octaves = 3, bool suppressScaleNonmaxima = true)
 * -- reference brisk/include/brisk/brisk-feature-detector.h:51-84. */
int brisk_agast_detector_create(brisk_ctx* ctx, int thresh, int octaves, int suppress_scale_nonmaxima,
                                brisk_detector** out);
/* brisk::ScaleSpaceFeatureDetector<brisk::HarrisScoreCalculator>(size_t octaves, double uniformityRadius,
 * double absoluteThreshold = 0, size_t maxNumKpt = SIZE_MAX) -- reference
 * brisk/include/brisk/scale-space-feature-detector.h:62-135.  max_kpts < 0 means unlimited.  The
 * detector ignores masks, as the reference's detectImpl does (:100-101). */
int brisk_harris_detector_create(brisk_ctx* ctx, int octaves, double uniformity_radius, double absolute_threshold,
                                 int64_t max_kpts, brisk_detector** out);

/* brisk::HarrisFeatureDetector(double radius) -- the legacy single-scale detector, reference
 * brisk/include/brisk/harris-feature-detector.h:51-82, brisk/src/harris-feature-detector.cc:56-409 (covariances, 3x3 binomial
 * smoothing of brisk/src/vectorized-filters.cc:54-123, response, 8-neighbour maxima >= 64, score-sorted uniformity
 * enforcement on a half-resolution occupancy map).  Used with brisk_detect() (the mask argument is ignored, as in the
 * reference); key points come in the order of the uniformity pass: x, y integral, size 10, angle -1, response = score,
 * octave 0.  The reference indexes its occupancy map with x as the row: images for which that leaves the map (landscape
 * shapes) are refused with BRISK_ERR_UNSUPPORTED -- the reference accesses memory out of bounds there. */
int brisk_harris_legacy_detector_create(brisk_ctx* ctx, double radius, brisk_detector** out);
void brisk_detector_destroy(brisk_detector* det);
/* Raw-corner capacity per frame (all layers); default scales with the image area. */
int brisk_detector_set_corner_capacity(brisk_detector* det, int corners_per_frame);

/* brisk::BriskDescriptorExtractor(bool rotationInvariant, bool scaleInvariant, int version,
 * float patternScale) and the pattern-file constructors -- reference
 * brisk/include/brisk/brisk-descriptor-extractor.h:54-202.  version: 1 = briskV1
 * (60 points, 64-byte descriptor), 2 = briskV2 (66 points, 48 bytes).
 * pattern_file: NULL for the built-in pattern. */
int brisk_extractor_create(brisk_ctx* ctx, int rotation_invariant, int scale_invariant, int version,
                           float pattern_scale, const char* pattern_file, brisk_extractor** out);
void brisk_extractor_destroy(brisk_extractor* ext);
int brisk_extractor_descriptor_size(const brisk_extractor* ext); /* descriptorSize(): 48 / 64 */
/* Pattern tables as built on the host (for inspection / parity tests); any output may be NULL.
 * counts = {points, short pairs, long pairs, descriptor bytes}. */
int brisk_extractor_pattern(const brisk_extractor* ext, int32_t counts[4], float* points_xys, float* scale_list,
                            uint32_t* size_list, uint32_t* short_pairs, int32_t* long_pairs);

/* detect(): BriskFeatureDetector::detectImpl -- reference brisk/src/brisk-feature-detector.cc:77-85.
 * imgs: n frames of w x h u8, row stride `stride` bytes, frame stride `frame_pitch` bytes.
 * masks: NULL or n masks with the same geometry (key points on zero mask pixels are removed).
 * kps: [n][cap] out, counts: [n] out (true count; > cap means truncated and BRISK_ERR_CAPACITY). */
int brisk_detect(brisk_ctx* ctx, brisk_detector* det, const uint8_t* imgs, int n, int w, int h, size_t stride,
                 size_t frame_pitch, const uint8_t* masks, brisk_keypoint* kps, int32_t* counts, int cap);

/* ComputeScale(): BriskFeatureDetector::ComputeScale -- reference brisk/src/brisk-feature-detector.cc:87-92, i.e.
 * BriskScaleSpace::GetKeypoints with caller-provided key points (brisk/src/brisk-scale-space.cc:104-124): every
 * provided point is mapped into every pyramid layer, goes through the scale-space checks there (no 2-D
 * non-maximum test) and yields one key point per layer that accepts it (x, y, size, response refined; octave =
 * layer; class_id copied).  A layer that keeps none of a frame's points (they must fall inside its 3-pixel
 * border after x / scale - offset) is detected on instead, with the threshold map's lower bound at 0, and its
 * corners go through the same checks (brisk/src/brisk-layer.cc:103-105).  kps_in: [n][cap_in], counts_in: [n]
 * (>= 1 each; with an empty vector the reference detects on all layers -- use brisk_detect); kps_out:
 * [n][cap_out], counts_out: [n] (true count; > cap_out means truncated and BRISK_ERR_CAPACITY). */
int brisk_compute_scale(brisk_ctx* ctx, brisk_detector* det, const uint8_t* imgs, int n, int w, int h, size_t stride,
                        size_t frame_pitch, const brisk_keypoint* kps_in, const int32_t* counts_in, int cap_in,
                        brisk_keypoint* kps_out, int32_t* counts_out, int cap_out);

/* ScaleSpaceFeatureDetector::detect() on a NON-EMPTY key-point vector ("use passed key points") -- reference
 * brisk/include/brisk/scale-space-feature-detector.h:103-108, brisk/include/brisk/internal/scale-space-layer-inl.h:
 * 198-208,370-428.  No image is read (the reference computes no scores in this mode; w, h size the occupancy map /
 * the buckets): the points with response > 1e6 are taken as (int score, uint16 x, uint16 y), sorted, filtered by the
 * uniformity enforcement (uniformityRadius > 0) or the bucketing, and returned unrefined (x, y truncated, size 12,
 * octave 0, class_id -1).  If no point of a frame passes the response test its vector is returned unchanged.
 * Only for octaves == 0 (a second layer indexes its occupancy map out of bounds in the reference). */
int brisk_harris_detect_passed(brisk_ctx* ctx, brisk_detector* det, int n, int w, int h, const brisk_keypoint* kps_in,
                               const int32_t* counts_in, int cap_in, brisk_keypoint* kps_out, int32_t* counts_out, int cap_out);

/* compute(): BriskDescriptorExtractor::computeImpl -- reference
 * brisk/src/brisk-descriptor-extractor.cc:589-599,612-778.  kps/counts are in/out: key points too
 * close to the border are removed (order kept), `angle` is written.  desc: [n][cap][descriptor_size]. */
int brisk_describe(brisk_ctx* ctx, brisk_extractor* ext, const uint8_t* imgs, int n, int w, int h, size_t stride,
                   size_t frame_pitch, brisk_keypoint* kps, int32_t* counts, int cap, uint8_t* desc);

/* detect() followed by compute() without leaving the device. */
int brisk_detect_describe(brisk_ctx* ctx, brisk_detector* det, brisk_extractor* ext, const uint8_t* imgs, int n, int w,
                          int h, size_t stride, size_t frame_pitch, const uint8_t* masks, brisk_keypoint* kps,
                          int32_t* counts, int cap, uint8_t* desc);

/* The same call for a caller that streams batches (the reference's callers run detect + compute once per camera frame,
 * brisk_ros_demo/src/livedemo.cc:458-504): returns once the batch is queued and all but its LAST chunk are back in the
 * host buffers; that chunk is collected by the next brisk_detect_describe_async call of the same shape -- after it has
 * queued its own first chunk, so that the upload of batch k+1 runs under the kernels of batch k and the download of
 * batch k under the kernels of batch k+1 -- or by brisk_sync(), or by any other call on the context.  The outputs of
 * batch k (and the inputs, which are read until then) must therefore stay untouched until the call after next returns
 * or brisk_sync() does: alternate between two sets of host buffers.  Errors that only show on the device (capacity,
 * unsupported data) are reported by brisk_sync() / the call that collects the chunk.  Device buffers, timing mode and
 * contexts with pipelining switched off fall back to the blocking call. */
int brisk_detect_describe_async(brisk_ctx* ctx, brisk_detector* det, brisk_extractor* ext, const uint8_t* imgs, int n, int w,
                                int h, size_t stride, size_t frame_pitch, const uint8_t* masks, brisk_keypoint* kps,
                                int32_t* counts, int cap, uint8_t* desc);

/* brisk::HarrisScoreCalculator::SetImage (InitializeScores = HarrisScoresSSE) and Get2dMaxima -- reference
 * brisk/include/brisk/harris-score-calculator.h:52-90, brisk/src/harris-score-calculator.cc:53-106, brisk/src/harris-scores.cc:
 * 53-279.  One 8-bit image.  scores (nullable): the int32 score map, h x w tightly packed, host memory.  maxima_sxy (nullable):
 * (score, x, y) triples of the 8-neighbour maxima >= abs_threshold in raster order, at most `cap`; *n_maxima receives the true
 * number (BRISK_ERR_CAPACITY when it exceeds cap). */
int brisk_harris_scores(brisk_ctx* ctx, const uint8_t* img, int w, int h, size_t stride, int abs_threshold, int32_t* scores,
                        int32_t* maxima_sxy, int cap, int32_t* n_maxima);

/* brisk::Halfsample16 / brisk::Twothirdsample16 -- reference brisk/src/image-down-sampling.cc:56-139, 394-548
 * (declared in brisk/include/brisk/internal/image-down-sampling.h:50-53).  CV_16UC1 images in host or device memory,
 * strides in bytes; dst is (h / 2) x (w / 2) resp. 2 (h / 3) x 2 (w / 3).  Bit-exact including the reference's quirks
 * (the doubled +1 on the lower-left pixel of Halfsample16, the signed saturation at 32767 of Twothirdsample16).  Images
 * narrower than one SSE block (16 resp. 12 columns), which the reference leaves unwritten, are refused. */
int brisk_halfsample16(brisk_ctx* ctx, const uint16_t* src, int w, int h, size_t src_stride, uint16_t* dst, size_t dst_stride);
int brisk_twothirdsample16(brisk_ctx* ctx, const uint16_t* src, int w, int h, size_t src_stride, uint16_t* dst, size_t dst_stride);

/* Stage dumps for parity tests: pyramid layers of ONE frame concatenated (tight rows), dims[2*i] =
 * cols, dims[2*i+1] = rows; integral image (h+1)x(w+1) int32; raw corners (x, y, score) of all
 * layers in detection order with layer_counts[n_layers]. */
int brisk_debug_pyramid(brisk_ctx* ctx, int octaves, const uint8_t* img, int w, int h, size_t stride, uint8_t* out,
                        int32_t* dims, int* n_layers);
int brisk_debug_integral(brisk_ctx* ctx, const uint8_t* img, int w, int h, size_t stride, int32_t* out);
int brisk_debug_corners(brisk_ctx* ctx, brisk_detector* det, const uint8_t* img, int w, int h, size_t stride,
                        int32_t* corners_xys, int cap, int32_t* layer_counts);

/* Dense FAST 9-16 / AGAST 5-8 score planes (w*h bytes each, threshold 1) of one image. */
int brisk_debug_scores(brisk_ctx* ctx, const uint8_t* img, int w, int h, size_t stride, uint8_t* out916, uint8_t* out58);
/* Corner-map (u16) and touch-map (u8) planes of all layers (tight rows, concatenated) after the
 * full detection of one image: internal NMS state, see ethzasl_brisk_b200/csrc/brisk_common.cuh. */
int brisk_debug_nms_state(brisk_ctx* ctx, brisk_detector* det, const uint8_t* img, int w, int h, size_t stride,
                          uint16_t* cm_out, uint8_t* bm_out);

/* Diagnostic: corners per layer (up to 12) that went through IsMax2D's tie path, frame 0 of the last detect call. */
int brisk_debug_nms_ties(brisk_ctx* ctx, int32_t* ties);

/* brisk::Hamming::operator()(a, b, size) -- reference brisk/include/brisk/internal/hamming.h:101-113,
 * hamming-inl.h:85-134: dist[i] = popcount(a[i] xor b[i]) over n descriptor pairs of desc_bytes each (rows desc_bytes
 * apart), counted over desc_bytes / 16 whole 128-bit words as the reference does (any size; trailing bytes are ignored). */
int brisk_hamming_distance(brisk_ctx* ctx, const uint8_t* a, const uint8_t* b, int64_t n, int desc_bytes,
                           int32_t* dist);

/* knnMatch(): BruteForceMatcher::commonKnnMatchImpl -- reference brisk/src/brute-force-matcher.cc:80-162
 * (one train collection, no mask).  idx/dist: [nq][k]; ties go to the lowest train index; missing
 * neighbours (nt < k) are -1.  desc_bytes: any multiple of 4 up to 496 (whole 128-bit words are compared, as brisk::Hamming
 * does); the extractors' widths (48 / 64 / 128) with k <= 8 run on the tuned kernels (tcgen05 for k = 2), everything else on a
 * general one that selects the k best in passes of eight. */
int brisk_hamming_knn(brisk_ctx* ctx, const uint8_t* query, int64_t nq, const uint8_t* train, int64_t nt,
                      int desc_bytes, int k, int32_t* idx, int32_t* dist);
/* std::sort of one query's match list by distance, entry for entry as libstdc++ (GCC 13) leaves it -- the reference sorts
 * every list it returns (brute-force-matcher.cc:160,210), and beyond 16 entries introsort permutes equal distances.
 * knnMatch lists of k > 16 need it (shorter ones come back in that order already); host arrays, no device work. */
int brisk_std_sort_matches(int64_t n, int32_t* train_idx, int32_t* img_idx, float* distance);

/* knnMatch() with masks: commonKnnMatchImpl's `isPossibleMatch` test -- reference brisk/src/brute-force-matcher.cc:
 * 118-119.  mask: [nq][nt] bytes, 0 = pair excluded (the per-image masks of a train collection side by side, as the
 * train rows are); NULL = no mask.  Queries with fewer than k allowed rows get -1 entries (the host classes turn
 * those into what the reference returns, see include/brisk/brisk.h). */
int brisk_hamming_knn_masked(brisk_ctx* ctx, const uint8_t* query, int64_t nq, const uint8_t* train, int64_t nt,
                             int desc_bytes, int k, const uint8_t* mask, int32_t* idx, int32_t* dist);
/* radiusMatch(): BruteForceMatcher::commonRadiusMatchImpl -- reference brisk/src/brute-force-matcher.cc:164-214.
 * Every train row with (float)distance < max_distance (and a non-zero mask byte, if a mask is given).  The
 * matches of query i are idx/dist[offsets[i] .. offsets[i+1]); offsets has nq + 1 entries.  sort = 0: train
 * order; sort = 1: the reference's order, i.e. std::sort by distance as libstdc++ performs it on the
 * train-ordered list (equal distances included).  If more than `capacity` matches exist the call
 * returns BRISK_ERR_CAPACITY with offsets filled in (offsets[nq] = capacity needed) and the first
 * `capacity` matches in train order. */
int brisk_hamming_radius(brisk_ctx* ctx, const uint8_t* query, int64_t nq, const uint8_t* train, int64_t nt, int desc_bytes,
                         float max_distance, const uint8_t* mask, int sort, int64_t* offsets, int32_t* idx, int32_t* dist,
                         int64_t capacity);
/* Sharded train set: local top-k as packed keys (dist << 32 | global train index), DEVICE buffer
 * keys[nq][k]; merge n_shards gathered key sets ([shard][nq][k], device) into idx/dist.  The exchange
 * between GPUs (NCCL all-gather of the keys) is done by the caller. */
int brisk_hamming_knn_keys(brisk_ctx* ctx, const uint8_t* query, int64_t nq, const uint8_t* train_shard, int64_t nt,
                           int desc_bytes, int k, int64_t global_train_offset, uint64_t* keys_dev);
int brisk_knn_merge_keys(brisk_ctx* ctx, const uint64_t* gathered_keys_dev, int n_shards, int64_t nq, int k,
                         int32_t* idx, int32_t* dist);

#ifdef __cplusplus
}
#endif
#endif /* BRISK_B200_H_ */
