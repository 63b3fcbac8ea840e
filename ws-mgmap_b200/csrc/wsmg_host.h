// Host-side helpers shared by the CUDA library (wsmg.cu) and the emulation (wsmg_emul.cpp):
// argument validation and the geometry constants, produced the way the reference's
// Python produces them (double arithmetic, one cast to fp32 at the point of use).
#pragma once
#include <math.h>

#include "../../include/wsmg.h"
#include "wsmg_math.h"

namespace wsmg {

inline int validate_dims(const wsmg_dims* d) {
  if (d == nullptr) return WSMG_E_NULL;
  if (d->bs <= 0 || d->C <= 0 || d->Hf <= 0 || d->Wf <= 0 || d->Hd <= 0 || d->Wd <= 0 || d->E <= 0 || d->G <= 0 ||
      !(d->resolution > 0.0))
    return WSMG_E_DIMS;
  if (d->Hd != d->Wd || d->Hf != d->Wf) return WSMG_E_DIMS;   // the reference assumes square frames (rgb_mapping.py:149-151,189)
  if (d->E > d->G) return WSMG_E_EGO_GT_GLOBAL;
  if (d->n_maps < d->bs) return WSMG_E_BATCH;
  if (d->C_in < 0) return WSMG_E_CHANNELS;
  if (d->feat_nhwc != 0 && (d->feat_nhwc != 1 || d->C % 4 != 0 || (d->C_in != 0 && d->C_in != d->C))) return WSMG_E_CHANNELS;
  if ((d->Hf * d->Wf) % 4 != 0) return WSMG_E_ALIGN;
  if ((long long)d->Hf * d->Wf > 63LL * 2048) return WSMG_E_DIMS;   // per-block flag words of the k_cells launch (wsmg_host.h MAX_FLAG_WORDS)
  if (d->E > 126 || d->G > 32768) return WSMG_E_DIMS;          // 16-bit fan codes; (E+2)/8 bands must fit the barrier array; row tables of 128
  return WSMG_OK;
}

inline Geo make_geo(const wsmg_dims* d) {
  Geo g;
  g.Cin = d->C_in > 0 ? d->C_in : d->C;
  g.feat_nhwc = d->feat_nhwc;
  g.E = d->E; g.G = d->G; g.C = d->C; g.Hf = d->Hf; g.Wf = d->Wf; g.Hd = d->Hd; g.Wd = d->Wd;
  const double cmin = -(double)d->G * d->resolution / 2;       // rgb_mapping.py:21
  const double cmax = (double)d->G * d->resolution / 2;        // rgb_mapping.py:22
  g.cmax = (float)cmax;
  g.cmin = (float)cmin;
  g.cell = (float)((cmax - cmin) / (double)d->G);              // rgb_mapping.py:98 == :146
  g.inv_cell = 1.0 / (double)g.cell;
  g.half = (float)((d->E - 1) / 2.0);                          // rgb_mapping.py:173
  const double t45 = tan(45.0 * (M_PI / 180.0));               // np.tan(np.deg2rad(fov/2)), fov = 90
  g.cx = (float)(d->Hd / 2.0);                                 // rgb_mapping.py:149
  g.cy = (float)(d->Wd / 2.0);
  g.fx = (float)((d->Hd / 2.0) / t45);                         // rgb_mapping.py:150
  g.fy = (float)((d->Wd / 2.0) / t45);
  g.ksub = (float)((double)d->Wd / (double)d->Wf);             // rgb_mapping.py:189
  g.half_e = (float)d->E / 2.0f;
  g.half_g = (float)d->G / 2.0f;
  g.gcenter = (float)(d->G / 2);                               // G//2
  g.paste_lo = d->G / 2 - d->E / 2;                            // rgb_mapping.py:42
  g.m_E = fd_magic(d->E); g.m_WW = fd_magic(d->E + 2); g.m_tiles = fd_magic((d->E + 7) >> 3);
  int ymax = d->E / 2;                                         // rint((E-1)/2 - a) <= ceil((E-1)/2) for a > 0
  g.fan_rows = ymax + 1 < d->E ? ymax + 1 : d->E;
  g.fan_cells = 0;
  for (int y = 0; y < g.fan_rows; ++y) g.fan_cells += fan_row_width(y, d->E);
  return g;
}

inline const char* error_string(int code) {
  switch (code) {
    case WSMG_OK: return "success";
    case WSMG_E_NULL: return "a required pointer is NULL";
    case WSMG_E_DIMS: return "non-positive or inconsistent dimension";
    case WSMG_E_EGO_GT_GLOBAL: return "egocentric map larger than the global map";
    case WSMG_E_CHANNELS: return "channel count not supported";
    case WSMG_E_SMEM: return "geometry needs more shared memory than one SM provides";
    case WSMG_E_ALIGN: return "pointer not 16-byte aligned or Hf*Wf not a multiple of 4";
    case WSMG_E_SCRATCH: return "scratch buffer too small";
    case WSMG_E_BATCH: return "bs larger than the map tensor's leading dimension";
    case WSMG_E_HOSTMEM: return "zero-copy needs page-locked, device-mapped host memory";
    default: return nullptr;
  }
}

// scratch layout: [bs * Hf*Wf] uint16 packed cell codes | [bs] uint32 env flags | [bs * E] int32 rotation column bounds |
// [bs * 4] fp32 rotation sines / cosines | [bs * 64] uint32 per-block flag words of the k_cells launch
inline size_t pad256(size_t n) { return (n + 255) & ~(size_t)255; }
inline size_t scratch_codes_bytes(const wsmg_dims* d) { return pad256((size_t)d->bs * d->Hf * d->Wf * sizeof(uint16_t)); }
inline size_t scratch_flags_bytes(const wsmg_dims* d) { return pad256((size_t)d->bs * sizeof(uint32_t)); }
inline size_t scratch_bounds_bytes(const wsmg_dims* d) { return pad256((size_t)d->bs * d->E * sizeof(int32_t)); }
inline size_t scratch_trig_bytes(const wsmg_dims* d) { return pad256((size_t)d->bs * 4 * sizeof(float)); }
constexpr int MAX_FLAG_WORDS = 64;       // k_cells blocks per env + 1 (validated: Hf*Wf <= 63 * 2048)
inline size_t scratch_blockflags_bytes(const wsmg_dims* d) { return pad256((size_t)d->bs * MAX_FLAG_WORDS * sizeof(uint32_t)); }
inline size_t scratch_bytes(const wsmg_dims* d) {
  return scratch_codes_bytes(d) + scratch_flags_bytes(d) + scratch_bounds_bytes(d) + scratch_trig_bytes(d) + scratch_blockflags_bytes(d);
}

// Views into a scratch buffer.
struct ScratchView { uint16_t* codes; uint32_t* flags; int32_t* bounds; float* env_trig; uint32_t* block_flags; };
inline ScratchView scratch_view(void* scratch, const wsmg_dims* d) {
  unsigned char* p = (unsigned char*)scratch;
  ScratchView v;
  v.codes = (uint16_t*)p; p += scratch_codes_bytes(d);
  v.flags = (uint32_t*)p; p += scratch_flags_bytes(d);
  v.bounds = (int32_t*)p; p += scratch_bounds_bytes(d);
  v.env_trig = (float*)p; p += scratch_trig_bytes(d);
  v.block_flags = (uint32_t*)p;
  return v;
}

inline void base_coords_host(float* out, int n) {
  for (int j = 0; j < n; ++j) out[j] = base_coord(j, n);
}

}  // namespace wsmg
