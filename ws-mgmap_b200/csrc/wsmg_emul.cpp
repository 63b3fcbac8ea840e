// TEST INFRASTRUCTURE ONLY -- serial host emulation of the CUDA kernels.
// Compiles the very same CTA body (wsmg_body.h) and pixel math (wsmg_math.h) with g++ and
// runs every "thread" in order (tid0 = 0, stride = 1), so the CPU test-suite can check the
// index arithmetic, the packed-fan layout, the band/ring schedule and the in-place hazards
// against the oracle without a GPU.  It cannot find races; the GPU parity tests do that.
// Not linked into libwsmg.so and not reachable from the product API.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

struct uint2 { unsigned x, y; };

#include "wsmg_body.h"
#include "wsmg_host.h"

using namespace wsmg;

extern "C" {

int wsmg_emul_unproject_index(const float* depth, int32_t* lin, uint8_t* invalid, uint16_t* codes, uint32_t* env_flags,
                              const wsmg_dims* d) {
  int rc = validate_dims(d);
  if (rc) return rc;
  const Geo g = make_geo(d);
  std::vector<int> rowoff(g.fan_rows + 1, 0);
  for (int y = 0; y < g.fan_rows; ++y) rowoff[y + 1] = rowoff[y] + fan_row_width(y, g.E);
  const int HW = g.Hf * g.Wf;
  for (int b = 0; b < d->bs; ++b)
    for (int t = 0; t < HW; ++t) {
      int i = t / g.Wf, j = t - i * g.Wf, x, y;
      // same split as k_cells: per-row / per-column pinhole terms, then the depth-dependent part
      const int r = sample_index(g, i), c = sample_index(g, j);
      bool ok = unproject_depth(g, depth[(size_t)b * g.Hd * g.Wd + (size_t)r * g.Wd + c], pinhole_xx(g, c), pinhole_yy(g, r), &x, &y);
      if (env_flags && !ok) env_flags[b] |= 1u;
      if (codes) {
        uint16_t code = CODE_INVALID;
        if (ok) {
          if (y < g.fan_rows && x >= fan_x_lo(y) && x <= fan_x_hi(y, g.E)) code = (uint16_t)(rowoff[y] + x - fan_x_lo(y));
          else code = CODE_OUTLIER;
        }
        codes[(size_t)b * HW + t] = code;
      }
      if (lin) lin[(size_t)b * HW + t] = y * g.E + x;
      if (invalid) invalid[(size_t)b * HW + t] = ok ? 0 : 1;
    }
  return 0;
}

// mode: 0 = whole step, 1 = scatter only (proj_out), 2 = registration only (proj_in)
int wsmg_emul_step(const float* feat, const float* depth, const float* gps, const float* compass, const float* mask,
                   float* gmap, float* ego_out, const float* trig, float* proj_out, const float* proj_in, int mode,
                   const wsmg_dims* d, uint16_t* ego_half, const int32_t* env_slots) {
  int rc = validate_dims(d);
  if (rc) return rc;
  const Geo g = make_geo(d);
  const SmemPlan sp = make_plan(g);
  const int HW = g.Hf * g.Wf;
  const char* force = getenv("WSMG_FORCE_GENERIC");
  const bool generic = force && force[0] == '1';
  std::vector<uint16_t> codes((size_t)d->bs * HW);
  std::vector<uint32_t> flags(d->bs, 0u);
  if (mode != 2) wsmg_emul_unproject_index(depth, nullptr, nullptr, codes.data(), flags.data(), d);
  if (mode != 1) {
    const size_t per_env = (size_t)g.G * g.G * g.C;
    for (int b = 0; b < d->bs; ++b) {
      float m = mask[b];
      if (m == 1.0f) continue;
      float* base = gmap + (size_t)(env_slots ? env_slots[b] : b) * per_env;
      for (size_t i = 0; i < per_env; ++i) base[i] = (m == 0.0f) ? 0.0f : base[i] * m;
    }
  }
  std::vector<int32_t> bounds((size_t)d->bs * g.E, g.E << 16);
  if (mode != 1) {                                 // what k_reset computes per env
    for (int b = 0; b < d->bs; ++b) {
      float cs, sn;
      if (trig) { cs = trig[4 * b + 0]; sn = trig[4 * b + 1]; }
      else { const float h = -compass[b]; sn = sinf(h); cs = cosf(h); }
      for (int t = 0; t < g.E; ++t) bounds[(size_t)b * g.E + t] = rot_row_bounds(g, cs, sn, t);
    }
  }
  FusedParams p{};
  p.row_bounds = bounds.data();
  p.feat = feat; p.codes = codes.data(); p.env_flags = flags.data(); p.gps = gps; p.compass = compass; p.trig = trig; p.gmap = gmap;
  p.ego = ego_out; p.proj_out = proj_out; p.proj_in = (mode == 2) ? proj_in : nullptr;
  p.stop_after_scatter = (mode == 1); p.bs = d->bs; p.g = g; p.sp = sp; p.ego_half = ego_half; p.env_slots = env_slots;
  p.n_maps = d->n_maps;
  std::vector<unsigned char> smem(sp.total + 128);
  unsigned char* sm = smem.data() + ((128 - ((uintptr_t)smem.data() & 127)) & 127);
  const int slabs = (g.C + SLAB - 1) / SLAB;
  for (int blk = 0; blk < d->bs * slabs; ++blk) {
    memset(sm, 0xCD, sp.total);     // poison: nothing may rely on zeroed shared memory
    const bool pool = g.Cin != g.C;
    if (pool && g.C % 4 == 0) fused_body<1, 0, 0, 0, true, false, FEAT_POOL>(p, blk, sm, 0);
    else if (pool) fused_body<1, 0, 0, 0, false, false, FEAT_POOL>(p, blk, sm, 0);
    else if (g.feat_nhwc) fused_body<1, 0, 0, 0, true, false, FEAT_NHWC>(p, blk, sm, 0);
    else if (g.C % 4 == 0 && g.E == 100 && g.G == 240 && HW == 224 * 224 && !generic) fused_body<1, 100, 240, 224 * 224, true, false, FEAT_NCHW>(p, blk, sm, 0);
    else if (g.C % 4 == 0) fused_body<1, 0, 0, 0, true, false, FEAT_NCHW>(p, blk, sm, 0);
    else fused_body<1, 0, 0, 0, false, false, FEAT_NCHW>(p, blk, sm, 0);
  }
  return 0;
}

// host mirror of k_semcrop (wsmg.cu)
int wsmg_emul_semantic_crop(const float* maps, const float* pose, const float* trig, const int32_t* map_index, int64_t* out,
                            int32_t bs, int32_t n_maps, int32_t S, int32_t half, int32_t origin) {
  const int side = 2 * half;
  for (int b = 0; b < bs; ++b) {
    float cs, sn;
    if (trig) { cs = trig[2 * b]; sn = trig[2 * b + 1]; }
    else { const float h = pose[3 * b + 2]; cs = cosf(h); sn = sinf(h); }
    const int m = map_index ? map_index[b] : b;
    for (int t = 0; t < side * side; ++t) {
      const int i = t / side, j = t - i * side;
      const int idx = semmap_source_index(origin - side + i, origin - side + j, S, cs, sn, pose[3 * b], pose[3 * b + 1]);
      out[(size_t)b * side * side + t] = (idx >= 0 && (unsigned)m < (unsigned)n_maps) ? (int64_t)maps[(size_t)m * S * S + idx] : 0;
    }
  }
  return 0;
}

int wsmg_emul_smem_bytes(const wsmg_dims* d) {
  if (validate_dims(d)) return -1;
  return make_plan(make_geo(d)).total;
}

int wsmg_emul_fan_cells(const wsmg_dims* d) {
  if (validate_dims(d)) return -1;
  return make_geo(d).fan_cells;
}

}  // extern "C"
