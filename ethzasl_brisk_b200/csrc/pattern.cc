// Host-side construction of the BRISK pattern tables.
//
// Mirrors what the reference computes once per extractor (brisk/src/
// brisk-descriptor-extractor.cc:65-178 generateKernel for BRISK 1.0, :180-291
// InitFromStream for the BRISK2 text pattern, :618-650 for the size->scale
// index mapping), with the same float/double promotions and the same libm
// (cos, sin, pow, log in double), because the descriptor bits depend on the
// rounded table entries.  The table is then uploaded once; kernels never
// recompute trigonometry.
#include "pattern.h"

#include <cmath>
#include <cstring>
#include <fstream>
#include <sstream>

#include "brisk_pattern_data.inc"
#include "describe_logic.cuh"

namespace briskb200 {
namespace {

constexpr unsigned kScales = 64, kRot = 1024;
const double kPi = 3.14159265358979323846;

float lb_scale_step() {
  static const float lb_scale = (float)(std::log((double)30.0f) / std::log(2.0));
  static const float step = lb_scale / (float)kScales;
  return step;
}

// brisk-descriptor-extractor.cc:636-646: scale index of a key-point size.
int scale_index(float size) {
  static const float ln2f = 0.693147180559945f;
  static const float lb_scalerange = (float)(std::log((double)30.0f) / (double)ln2f);
  static const float basic_size06 = (float)((double)12.0f * 0.6);
  const double v = (double)((float)kScales / lb_scalerange) * (std::log((double)(size / basic_size06)) / (double)ln2f) + 0.5;
  int idx;
  if (!(v > -2147483648.0 && v < 2147483648.0)) idx = -2147483647 - 1;  // cvttsd2si out-of-range / NaN result
  else idx = (int)v;
  if (idx < 0) idx = 0;
  if (idx > (int)kScales - 1) idx = kScales - 1;
  return idx;
}

void finish(PatternHost* p) {
  const int ns = (int)p->short_pairs.size() / 2;
  p->sample_consts.resize((size_t)kScales * p->n_points * 2);
  for (unsigned s = 0; s < kScales; ++s)
    for (int i = 0; i < p->n_points; ++i) {
      const float sigma = p->points[(((size_t)s * kRot) * p->n_points + i) * 3 + 2];  // same for every rotation
      sampling_constants(sigma, &p->sample_consts[((size_t)s * p->n_points + i) * 2], &p->sample_consts[((size_t)s * p->n_points + i) * 2 + 1]);
    }
  p->desc_bytes = (int)std::ceil((float)ns / 128.0) * 16;
  // smallest float size that reaches each scale index (monotone in size)
  p->scale_breaks[0] = 0.0f;
  for (int s = 1; s < (int)kScales; ++s) {
    uint32_t lo = 0x00800000u /* FLT_MIN */, hi = 0x7f7fffffu /* FLT_MAX */;
    while (lo < hi) {
      const uint32_t mid = lo + (hi - lo) / 2;
      float f; std::memcpy(&f, &mid, 4);
      if (scale_index(f) >= s) hi = mid; else lo = mid + 1;
    }
    float f; std::memcpy(&f, &lo, 4);
    p->scale_breaks[s] = f;
  }
  // brisk-descriptor-extractor.cc:630-635
  static const float ln2f = 0.693147180559945f;
  static const float lb_scalerange = (float)(std::log((double)30.0f) / (double)ln2f);
  static const float basic_size06 = (float)((double)12.0f * 0.6);
  int b = (int)((double)((float)kScales / lb_scalerange) * (std::log(1.45 * (double)12.0f / (double)basic_size06) / (double)ln2f) + 0.5);
  p->basic_scale = b < 0 ? 0 : b;
}

std::string init_from_tokens(std::istream& in, float pattern_scale, PatternHost* p) {
  unsigned n = 0;
  if (!(in >> n) || n == 0 || n > 96) return "bad pattern: point count";
  p->n_points = (int)n;
  std::vector<float> ux(n), uy(n), sg(n);
  for (unsigned i = 0; i < n; ++i) {
    if (!(in >> ux[i] >> uy[i] >> sg[i])) return "bad pattern: points";
    ux[i] *= pattern_scale; uy[i] *= pattern_scale; sg[i] *= pattern_scale;
  }
  p->points.resize((size_t)kScales * kRot * n * 3);
  const float sigma_scale = 1.3f;
  float* it = p->points.data();
  for (unsigned s = 0; s < kScales; ++s) {
    p->scale_list[s] = (float)std::pow(2.0, (double)((float)s * lb_scale_step()));
    p->size_list[s] = 0;
    for (unsigned rot = 0; rot < kRot; ++rot) {
      const double theta = (double)rot * 2 * kPi / (double)kRot;
      const double c = std::cos(theta), sn = std::sin(theta);
      for (unsigned i = 0; i < n; ++i, it += 3) {
        it[0] = (float)((double)p->scale_list[s] * ((double)ux[i] * c - (double)uy[i] * sn));
        it[1] = (float)((double)p->scale_list[s] * ((double)ux[i] * sn + (double)uy[i] * c));
        it[2] = sigma_scale * p->scale_list[s] * sg[i];
        const unsigned size = (unsigned)(std::ceil(std::sqrt((double)(it[0] * it[0] + it[1] * it[1])) + (double)it[2]) + 1);
        if (p->size_list[s] < size) p->size_list[s] = size;
      }
    }
  }
  unsigned ns = 0;
  if (!(in >> ns)) return "bad pattern: short pair count";
  p->short_pairs.resize((size_t)ns * 2);
  for (unsigned k = 0; k < ns; ++k) {
    unsigned i, j;
    if (!(in >> i >> j) || i >= n || j >= n) return "bad pattern: short pairs";
    p->short_pairs[2 * k] = (unsigned short)i; p->short_pairs[2 * k + 1] = (unsigned short)j;
  }
  unsigned nl = 0;
  if (!(in >> nl)) return "bad pattern: long pair count";
  p->long_pairs.resize((size_t)nl * 4);
  for (unsigned k = 0; k < nl; ++k) {
    unsigned i, j;
    if (!(in >> i >> j) || i >= n || j >= n) return "bad pattern: long pairs";
    const float dx = ux[j] - ux[i], dy = uy[j] - uy[i];
    const float nsq = dx * dx + dy * dy;
    p->long_pairs[4 * k] = (int)i; p->long_pairs[4 * k + 1] = (int)j;
    p->long_pairs[4 * k + 2] = (int)((double)(dx / nsq) * 2048.0 + 0.5);
    p->long_pairs[4 * k + 3] = (int)((double)(dy / nsq) * 2048.0 + 0.5);
  }
  // the reference CHECKs noShortPairs == 384 on this path (:286)
  if (ns != 384) return "pattern must define exactly 384 short pairs (reference kDescriptorLength)";
  finish(p);
  return "";
}

std::string init_v1(float pattern_scale, PatternHost* p) {
  const double f = 0.85 * (double)pattern_scale;
  const float radius[5] = {(float)(f * 0), (float)(f * 2.9), (float)(f * 4.9), (float)(f * 7.4), (float)(f * 10.8)};
  const int number[5] = {1, 10, 14, 15, 20};
  const float d_max = 5.85f, d_min = 8.2f;
  int n = 0;
  for (int r = 0; r < 5; ++r) n += number[r];
  p->n_points = n;
  p->points.resize((size_t)kScales * kRot * n * 3);
  const float sigma_scale = 1.3f;
  float* it = p->points.data();
  for (unsigned s = 0; s < kScales; ++s) {
    p->scale_list[s] = (float)std::pow(2.0, (double)((float)s * lb_scale_step()));
    p->size_list[s] = 0;
    for (unsigned rot = 0; rot < kRot; ++rot) {
      const double theta = (double)rot * 2 * kPi / (double)kRot;
      for (int ring = 0; ring < 5; ++ring)
        for (int num = 0; num < number[ring]; ++num, it += 3) {
          const double alpha = ((double)num) * 2 * kPi / (double)number[ring];
          const float sr = p->scale_list[s] * radius[ring];
          it[0] = (float)((double)sr * std::cos(alpha + theta));
          it[1] = (float)((double)sr * std::sin(alpha + theta));
          if (ring == 0) it[2] = (float)((double)(sigma_scale * p->scale_list[s]) * 0.5);
          else it[2] = (float)((double)(sigma_scale * p->scale_list[s]) * (double)radius[ring] * std::sin(kPi / number[ring]));
          const unsigned size = (unsigned)(std::ceil(sr + it[2]) + 1);
          if (p->size_list[s] < size) p->size_list[s] = size;
        }
    }
  }
  const float dmin_sq = d_min * d_min, dmax_sq = d_max * d_max;
  const float* base = p->points.data();  // scale 0, rotation 0
  for (int i = 1; i < n; ++i)
    for (int j = 0; j < i; ++j) {
      const float dx = base[3 * j] - base[3 * i], dy = base[3 * j + 1] - base[3 * i + 1];
      const float nsq = dx * dx + dy * dy;
      if (nsq > dmin_sq) {
        p->long_pairs.push_back(i); p->long_pairs.push_back(j);
        p->long_pairs.push_back((int)((double)(dx / nsq) * 2048.0 + 0.5));
        p->long_pairs.push_back((int)((double)(dy / nsq) * 2048.0 + 0.5));
      }
      if (nsq < dmax_sq) { p->short_pairs.push_back((unsigned short)i); p->short_pairs.push_back((unsigned short)j); }
    }
  finish(p);
  return "";
}

}  // namespace

std::string build_pattern(int version, float pattern_scale, const char* pattern_file, PatternHost* out) {
  *out = PatternHost();
  if (pattern_file && pattern_file[0]) {
    std::ifstream f(pattern_file);
    if (!f.is_open()) return std::string("cannot open pattern file ") + pattern_file;
    return init_from_tokens(f, pattern_scale, out);
  }
  if (version == 2) {
    // re-create the reference's token stream from the embedded tables; floats
    // are printed with enough digits to round-trip exactly.
    std::ostringstream os;
    os.precision(9);
    os << kBrisk2NumPoints << "\n";
    for (int i = 0; i < 3 * kBrisk2NumPoints; ++i) {
      float v; std::memcpy(&v, &kBrisk2PointBits[i], 4);
      os << v << " ";
    }
    os << "\n" << kBrisk2NumShortPairs << "\n";
    for (int i = 0; i < 2 * kBrisk2NumShortPairs; ++i) os << (int)kBrisk2ShortPairs[i] << " ";
    os << "\n" << kBrisk2NumLongPairs << "\n";
    for (int i = 0; i < 2 * kBrisk2NumLongPairs; ++i) os << (int)kBrisk2LongPairs[i] << " ";
    std::istringstream is(os.str());
    return init_from_tokens(is, pattern_scale, out);
  }
  if (version == 1) return init_v1(pattern_scale, out);
  return "only version 1 (briskV1) or 2 (briskV2) supported";
}

}  // namespace briskb200
