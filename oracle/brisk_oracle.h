/* ORACLE (TEST INFRASTRUCTURE ONLY -- never linked into or called by the product).
 *
 * CPU restatement of the reference's BRISK hot path in closed-form scalar
 * arithmetic.  Every function cites the reference file:line it restates.
 * Pinned against (a) the reference's golden fixtures
 * (tests/golden/brisk_verification.npz, extracted from
 * brisk/src/test/test_data/brisk_verification_{ast,harris}.set) and (b) the
 * unmodified reference compiled into oracle/_ref (tests/test_oracle_golden.py, tests/test_host_logic.py).
 */
#ifndef BRISK_ORACLE_H_
#define BRISK_ORACLE_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  float x, y, size, angle, response;
  int32_t octave, class_id;
} orc_keypoint; /* == cv::KeyPoint, 28 bytes */

/* image-down-sampling.cc:142-392 / :550-787 */
int orc_sort_desc_by_response(const float* response, int32_t* perm, int n);
void orc_halfsample8(const uint8_t* src, int w, int h, uint8_t* dst);
void orc_twothirdsample8(const uint8_t* src, int w, int h, uint8_t* dst);
/* brisk-layer.cc:278-598 */
void orc_thrmap(const uint8_t* img, int w, int h, uint8_t* thr);
/* oast9-16-nms.cc:39-1976, agast5-8-nms.cc:39-358: dense score maps at threshold 1
 * through the lazy accessors' border rules (brisk-layer.cc:118-145) */
void orc_dense_scores(const uint8_t* img, int w, int h, uint8_t* out916, uint8_t* out58);
/* oast9-16.cc:43-1859 + brisk-layer.cc:99-117: raw corners (x, y, score) */
int orc_layer_corners(const uint8_t* img, int w, int h, int thresh, int lower, int32_t* xys, int cap);
/* brisk-scale-space.cc:64-90 */
int orc_pyramid(const uint8_t* img, int w, int h, int octaves, uint8_t* out, int32_t* dims, float* scale_offset);
/* brisk-feature-detector.cc:77-85 (BriskFeatureDetector::detectImpl) */
int orc_agast_detect(const uint8_t* img, int w, int h, int thresh, int octaves, int suppress,
                     const uint8_t* mask, orc_keypoint* out, int cap);
/* brisk-feature-detector.cc:87-92 (ComputeScale) = brisk-scale-space.cc:92-287 with provided key points */
int orc_compute_scale(const uint8_t* img, int w, int h, int thresh, int octaves, int suppress, const orc_keypoint* in,
                      int n_in, orc_keypoint* out, int cap);
/* debug aid: final lazy score cache of every layer (brisk-layer.cc:118-132) */
int orc_agast_cache_dump(const uint8_t* img, int w, int h, int thresh, int octaves, uint8_t* out);
/* integral-image.h:56-161 */
void orc_integral8(const uint8_t* img, int w, int h, int32_t* out);
/* brisk-descriptor-extractor.cc:612-778 */
int orc_describe(const uint8_t* img, int w, int h, orc_keypoint* kps, int n, int rot, int scale, int version,
                 float pattern_scale, uint8_t* desc, int32_t* desc_bytes);
/* brisk-descriptor-extractor.cc:65-291: pattern tables */
int orc_pattern_dump(int version, float pattern_scale, int32_t* counts, float* points_xys, float* scale_list,
                     uint32_t* size_list, uint32_t* short_pairs, int32_t* long_pairs);
/* harris-scores.cc:53-279 */
void orc_harris_scores(const uint8_t* img, int w, int h, int32_t* out);
/* harris-score-calculator.cc:57-106 */
int orc_harris_maxima(const uint8_t* img, int w, int h, int abs_thr, int32_t* out_sxy, int cap);
/* scale-space-feature-detector.h:100-128 */
int orc_harris_detect(const uint8_t* img, int w, int h, int octaves, double radius, double abs_thr, int64_t max_kpt,
                      orc_keypoint* out, int cap);
/* scale-space-feature-detector.h:100-128 with passed key points (one layer, octaves == 0) */
int orc_harris_detect_passed(int w, int h, double radius, int64_t max_kpt, const orc_keypoint* in, int n_in,
                             orc_keypoint* out, int cap);
/* hamming-inl.h:85-134 */
int orc_hamming(const uint8_t* a, const uint8_t* b, int nbytes);
/* brute-force-matcher.cc:80-162 (single train image, no mask) */
int orc_knn(const uint8_t* q, int64_t nq, const uint8_t* t, int64_t nt, int nbytes, int k, int32_t* idx,
            int32_t* dist);

/* brute-force-matcher.cc:160,210: std::sort of the DMatch list of one query (by distance, libstdc++ order) */
int orc_sort_matches(int32_t* train_idx, int32_t* img_idx, float* distance, int n);
/* all pairwise Hamming distances, out[nq][nt] */
int orc_hamming_matrix(const uint8_t* q, int64_t nq, const uint8_t* t, int64_t nt, int nbytes, int32_t* out);

#ifdef __cplusplus
}
#endif
#endif /* BRISK_ORACLE_H_ */
