// Host-side construction of the BRISK sampling-pattern tables (the look-up
// table the reference builds in BriskDescriptorExtractor's constructors).
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

namespace briskb200 {

struct PatternHost {
  int n_points = 0;
  std::vector<float> points;            // [64][1024][n_points][3] x, y, sigma
  float scale_list[64];
  unsigned int size_list[64];
  std::vector<unsigned short> short_pairs;  // [n][2] i, j
  std::vector<int> long_pairs;              // [n][4] i, j, weighted_dx, weighted_dy
  std::vector<int> sample_consts;       // [64][n_points][2] scaling, scaling2 (describe_logic.cuh)
  float scale_breaks[64];               // smallest key-point size mapping to scale index s
  int basic_scale = 0;                  // scale index used when scale invariance is off
  int desc_bytes = 0;
};

// version 2: default BRISK2 pattern (or `pattern_file` in the reference's .ptn
// text format); version 1: legacy 60-point BRISK ring pattern.  Returns an
// empty string on success, else an error message.
std::string build_pattern(int version, float pattern_scale, const char* pattern_file, PatternHost* out);

}  // namespace briskb200
