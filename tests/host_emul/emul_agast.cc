// TEST HARNESS: compiles the product's __host__ __device__ NMS / refinement
// logic (ethzasl_brisk_b200/csrc/*.cuh) for the CPU and runs the kernels'
// phases serially, so the data-parallel reformulation of the order-dependent
// reference algorithm can be checked against the oracle without a GPU.  This
// is a test of product device code, not a CPU fallback: nothing in the product
// links it.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define BRISK_TILE_CHECK 1  // count score-tile accesses outside the filled rectangle (must stay 0)
#include "../../ethzasl_brisk_b200/csrc/nms_logic.cuh"

using namespace briskb200;

namespace {
struct HostLayer {
  int w, h, pitch;
  std::vector<uint8_t> img, bm;
  std::vector<uint16_t> cm;
  float scale, offset;
  std::vector<int> cx, cy;  // raster-ordered corners
};
}  // namespace

namespace {
// pyramid as the pyramid kernel builds it (brisk_math.cuh sampling formulas)
std::vector<HostLayer> build_layers(const uint8_t* image, int w, int h, int octaves) {
  const int n = octaves == 0 ? 1 : 2 * octaves;
  std::vector<HostLayer> H(n);
  auto alloc = [](HostLayer& l, int w_, int h_) {
    l.w = w_; l.h = h_; l.pitch = (w_ + 15) / 16 * 16;
    l.img.assign((size_t)l.pitch * h_, 0); l.bm.assign((size_t)l.pitch * h_, 0); l.cm.assign((size_t)l.pitch * h_, 0);
  };
  alloc(H[0], w, h);
  for (int y = 0; y < h; ++y) memcpy(&H[0].img[(size_t)y * H[0].pitch], image + (size_t)y * w, w);
  H[0].scale = 1.0f; H[0].offset = 0.0f;
  for (int i = 1; i < n; ++i) {
    const HostLayer& s = (i == 1) ? H[0] : H[i - 2];
    HostLayer& d = H[i];
    if (i == 1) {
      alloc(d, 2 * (s.w / 3), 2 * (s.h / 3));
      for (int R = 0; R < s.h / 3; ++R)
        for (int T = 0; T < s.w / 3; ++T) {
          int p[9], o[4];
          for (int k = 0; k < 9; ++k) p[k] = s.img[(size_t)(3 * R + k / 3) * s.pitch + 3 * T + k % 3];
          twothird_block(p, T, s.w, o);
          d.img[(size_t)(2 * R) * d.pitch + 2 * T] = (uint8_t)o[0]; d.img[(size_t)(2 * R) * d.pitch + 2 * T + 1] = (uint8_t)o[1];
          d.img[(size_t)(2 * R + 1) * d.pitch + 2 * T] = (uint8_t)o[2]; d.img[(size_t)(2 * R + 1) * d.pitch + 2 * T + 1] = (uint8_t)o[3];
        }
      d.scale = (float)(s.scale * 1.5);
    } else {
      alloc(d, s.w / 2, s.h / 2);
      for (int r = 0; r < d.h; ++r)
        for (int c = 0; c < d.w; ++c) {
          const uint8_t* a = &s.img[(size_t)(2 * r) * s.pitch + 2 * c];
          d.img[(size_t)r * d.pitch + c] = (uint8_t)halfsample_px(a[0], a[1], a[s.pitch], a[s.pitch + 1], c, s.w);
        }
      d.scale = s.scale * 2;
    }
    d.offset = (float)(0.5 * d.scale - 0.5);
  }
  return H;
}
}  // namespace

extern "C" int emul_agast_detect_ex(const uint8_t* image, int w, int h, int thresh, int octaves, KeyPoint* out, int cap, uint16_t* cm_out, uint8_t* bm_out) {
  const int n = octaves == 0 ? 1 : 2 * octaves;
  std::vector<HostLayer> H = build_layers(image, w, h, octaves);
  std::vector<LayerView> V(n);
  for (int i = 0; i < n; ++i) {
    HostLayer& l = H[i];
    V[i] = LayerView{l.img.data(), l.cm.data(), l.bm.data(), l.w, l.h, l.pitch, l.scale, l.offset};
    // detect kernel: threshold map + segment test -> corner map, raster order
    for (int y = 3; y < l.h - 3; ++y)
      for (int x = 3; x < l.w - 3; ++x) {
        const int T = thrmap_px(l.img.data(), l.pitch, x, y);
        if (agast_is_corner(l.img.data(), l.pitch, x, y, T, thresh)) {
          if (T <= 2) return -400;  // corner_score_check_kernel: the closed form needs sticky corner scores (thresh < 20 only)
          l.cm[(size_t)y * l.pitch + x] = (uint16_t)T;
          l.cx.push_back(x); l.cy.push_back(y);
        }
      }
  }
  int violations = 0;
  // phase 1 + 2 (parallel on the GPU)
  std::vector<std::vector<uint8_t>> fwin(n);
  std::vector<std::vector<unsigned long long>> tiles(n);
  std::vector<std::vector<CheckResult>> chk(n);
  // nms_windows_kernel: the 5x5 score window, IsMax2D's comparisons and, for the corners that pass them, the score
  // tiles of the two neighbouring layers
  for (int i = 0; i < n; ++i) {
    const size_t nc = H[i].cx.size();
    fwin[i].assign(nc * 25, 0);
    tiles[i].assign(nc * 4, 0);
    chk[i].resize(nc);
    for (size_t k = 0; k < nc; ++k) nms_prefix(V[i], H[i].cx[k], H[i].cy[k], &fwin[i][k * 25]);
  }
  for (int i = 0; i < n; ++i)
    for (size_t k = 0; k < H[i].cx.size(); ++k) {
      const uint16_t e = H[i].cm[(size_t)H[i].cy[k] * H[i].pitch + H[i].cx[k]];
      if ((e & kCmDecided) && !(e & kCmAccept)) continue;
      uint8_t t[32];
      side_tiles(V[i > 0 ? i - 1 : 0], V[i], V[i + 1 < n ? i + 1 : i], n, i, H[i].cx[k], H[i].cy[k], t);
      memcpy(&tiles[i][k * 4], t, 32);
    }
  // nms_checks_kernel: the scans on the precomputed tiles
  for (int i = 0; i < n; ++i)
    for (size_t k = 0; k < H[i].cx.size(); ++k) {
      uint16_t& e = H[i].cm[(size_t)H[i].cy[k] * H[i].pitch + H[i].cx[k]];
      if ((e & kCmDecided) && !(e & kCmAccept)) continue;
      const int v0 = violations;
      if (nms_checks(V.data(), n, i, H[i].cx[k], H[i].cy[k], &chk[i][k], &violations, &tiles[i][k * 4])) e |= kCmChecks;
      if (violations != v0 && getenv("EMUL_DEBUG")) fprintf(stderr, "tile violation: layer %d corner (%d,%d) +%d\n", i, H[i].cx[k], H[i].cy[k], violations - v0);  // chk kept either way: footprint
    }
  if (violations) return -200;  // a scan looked outside its score tile
  // phase 3 + 4 (per frame: layers in order).
  for (int i = 0; i < n; ++i) {
    const int mode = n == 1 ? kModeSingle : (i == n - 1 ? kModeLast : kModeMid);
    // Passes as the GPU runs them: every undecided corner is evaluated against the SAME snapshot of the
    // corner map (a parallel pass), then the verdicts are published.
    for (int pass = 0;; ++pass) {
      std::vector<std::pair<size_t, int>> verdicts;
      size_t left = 0;
      for (size_t kk = H[i].cx.size(); kk-- > 0;) {
        const uint16_t e = H[i].cm[(size_t)H[i].cy[kk] * H[i].pitch + H[i].cx[kk]];
        if (e & kCmDecided) continue;
        uint16_t scratch[64];
        const int verdict = nms_tie_decide(V[i], mode, H[i].cx[kk], H[i].cy[kk], &fwin[i][kk * 25], scratch, 1);
        if (verdict < 0) { ++left; continue; }
        verdicts.emplace_back(kk, verdict);
      }
      for (auto& kv : verdicts) H[i].cm[(size_t)H[i].cy[kv.first] * H[i].pitch + H[i].cx[kv.first]] |= (uint16_t)(kCmDecided | (kv.second ? kCmAccept : 0));
      if (getenv("EMUL_DEBUG")) fprintf(stderr, "layer %d pass %d: decided %zu, left %zu\n", i, pass, verdicts.size(), left);
      if (!left) break;
      if (verdicts.empty()) return -100;  // no progress: must never happen
    }
    if (mode == kModeMid)
      for (size_t k = 0; k < H[i].cx.size(); ++k) {
        const uint16_t e = H[i].cm[(size_t)H[i].cy[k] * H[i].pitch + H[i].cx[k]];
        if (e & kCmAccept) mark_above(V.data(), i, H[i].cx[k], H[i].cy[k], chk[i][k]);
      }
  }
  // phase 5 + ordered compaction
  int total = 0;
  for (int i = 0; i < n; ++i)
    for (size_t k = 0; k < H[i].cx.size(); ++k) {
      const uint16_t e = H[i].cm[(size_t)H[i].cy[k] * H[i].pitch + H[i].cx[k]];
      if (!(e & kCmAccept) || !(e & kCmChecks)) continue;
      KeyPoint kp;
      if (!refine_emit(V.data(), n, i, H[i].cx[k], H[i].cy[k], chk[i][k], &kp, &fwin[i][k * 25])) continue;
      if (total < cap) out[total] = kp;
      ++total;
    }
  if (cm_out || bm_out) {
    size_t off = 0;
    for (int i = 0; i < n; ++i) {
      for (int y = 0; y < H[i].h; ++y)
        for (int x = 0; x < H[i].w; ++x) {
          if (cm_out) cm_out[off + (size_t)y * H[i].w + x] = H[i].cm[(size_t)y * H[i].pitch + x];
          if (bm_out) bm_out[off + (size_t)y * H[i].w + x] = H[i].bm[(size_t)y * H[i].pitch + x];
        }
      off += (size_t)H[i].w * H[i].h;
    }
  }
  return total;
}

extern "C" int emul_agast_detect(const uint8_t* image, int w, int h, int thresh, int octaves, KeyPoint* out, int cap) {
  return emul_agast_detect_ex(image, w, h, thresh, octaves, out, cap, nullptr, nullptr);
}

// BriskFeatureDetector::ComputeScale (provided key points): the passes of the GPU path (nms.cu, provided_*)
// run serially.  A layer that keeps no point is detected on with the threshold map's lower bound at 0.
extern "C" int emul_compute_scale(const uint8_t* image, int w, int h, int thresh, int octaves, const KeyPoint* in, int n_in, KeyPoint* out, int cap) {
  const int n = octaves == 0 ? 1 : 2 * octaves;
  std::vector<HostLayer> H = build_layers(image, w, h, octaves);
  std::vector<LayerView> V(n);
  for (int i = 0; i < n; ++i) V[i] = LayerView{H[i].img.data(), H[i].cm.data(), H[i].bm.data(), H[i].w, H[i].h, H[i].pitch, H[i].scale, H[i].offset};
  // pass 0: points kept per layer
  std::vector<int> kept(n, 0);
  bool fallback = false;
  for (int i = 0; i < n; ++i) {
    for (int k = 0; k < n_in; ++k) {
      float px, py;
      kept[i] += provided_to_layer(V[i], in[k].x, in[k].y, &px, &py);
    }
    fallback |= kept[i] == 0;
  }
  if (fallback) {
    // detect kernel (lower = 0) on every layer, then provided_clear_kernel
    for (int i = 0; i < n; ++i) {
      HostLayer& l = H[i];
      for (int y = 3; y < l.h - 3; ++y)
        for (int x = 3; x < l.w - 3; ++x) {
          const int T = thrmap_px(l.img.data(), l.pitch, x, y);
          if (agast_is_corner(l.img.data(), l.pitch, x, y, T, thresh, 0)) {
            l.cm[(size_t)y * l.pitch + x] = (uint16_t)T;
            l.cx.push_back(x); l.cy.push_back(y);
          }
        }
      for (auto& v : l.cm)
        if (v && (kept[i] > 0 || (v & kCmT) <= 2)) v = 0;
    }
  }
  for (int pass = 1; pass <= 2; ++pass)
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < n_in; ++k) {
        float px, py;
        if (!provided_to_layer(V[i], in[k].x, in[k].y, &px, &py)) continue;
        if (pass == 1) provided_touch(V[i], px, py); else provided_stamp(V[i], px, py);
      }
  int total = 0, violations = 0;
  for (int i = 0; i < n; ++i) {
    const LayerView &below = V[i > 0 ? i - 1 : 0], &above = V[i + 1 < n ? i + 1 : i];
    if (kept[i] > 0) {
      for (int k = 0; k < n_in; ++k) {
        float px, py;
        if (!provided_to_layer(V[i], in[k].x, in[k].y, &px, &py)) continue;
        KeyPoint kp;
        kp.class_id = in[k].class_id;
        if (!provided_refine(below, V[i], above, n, i, px, py, &kp, &violations)) continue;
        if (total < cap) out[total] = kp;
        ++total;
      }
    } else {
      for (size_t k = 0; k < H[i].cx.size(); ++k) {
        KeyPoint kp;
        kp.class_id = -1;
        if (!provided_refine(below, V[i], above, n, i, (float)H[i].cx[k], (float)H[i].cy[k], &kp, &violations)) continue;
        if (total < cap) out[total] = kp;
        ++total;
      }
    }
  }
  if (violations) return -200;  // a scan looked outside its score tile
  return total;
}
