// Shared definitions for the B200 BRISK kernels: layer geometry, device-side
// layout of a batch, and small helpers.  sm_100a only.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BRISK_HD __host__ __device__ __forceinline__
#define BRISK_D __device__ __forceinline__
#else
#define BRISK_HD inline
#define BRISK_D inline
#endif

namespace briskb200 {

constexpr int kMaxLayers = 12;  // fused pyramid kernel: up to 6 octaves

// AGAST constants of the reference (brisk-scale-space.cc:45-51).
constexpr int kLowerThreshold = 10;
constexpr int kUpperThreshold = 230;
constexpr int kDropThreshold = 5;
constexpr int kMaxThreshold = 1;
constexpr int kMinDrop = 15;

// One pyramid layer inside a frame's block.  All per-pixel planes of a layer
// (image u8, corner map u16, touch map u8) share width/height/pitch (in
// ELEMENTS) and differ only in element size and base pointer.
struct LayerGeom {
  int w, h;        // layer size in pixels
  int pitch;       // row pitch in elements (multiple of 16)
  int pad_;
  long long off;   // element offset of the layer inside one frame's plane block
  float scale;     // reference brisk-layer.cc:62-63,80-86
  float offset;
};

struct PyramidGeom {
  int n_layers;
  int w0, h0;
  int pad_;
  long long frame_elems;  // elements per frame in a plane block (sum of pitch*h, 256-aligned)
  LayerGeom L[kMaxLayers];
};

// cv::KeyPoint-compatible record (reference uses cv::KeyPoint; 28 bytes).
struct KeyPoint {
  float x, y, size, angle, response;
  int32_t octave, class_id;
};

// Corner-map entry (u16 per pixel of every layer), written by the detect
// kernel and updated by the NMS kernels.  0 == not a corner.
//   bits 0..7   threshold-map value T at the corner (== its stored score,
//               reference brisk-layer.cc:110-116)
//   bits 8..11  number of IsMax2D neighbour look-ups the corner performs (1..8)
//   bit  12     corner passed the 8-neighbour test and has >= 1 tying neighbour
//   bit  13     scale-space checks above/below passed (Refine3D reaches its
//               own-layer 3x3 patch)
//   bit  14     accepted by IsMax2D (final)
//   bit  15     decided (set once bit 14 is final)
constexpr uint16_t kCmT = 0x00ff;
constexpr int kCmCallsShift = 8;
constexpr uint16_t kCmCalls = 0x0f00;
constexpr uint16_t kCmTie = 0x1000;
constexpr uint16_t kCmChecks = 0x2000;
constexpr uint16_t kCmAccept = 0x4000;
constexpr uint16_t kCmDecided = 0x8000;

}  // namespace briskb200
