// Body of the fused map-update CTA, written once and compiled twice:
//   * by nvcc as the device code of k_fused (wsmg.cu), one CTA per (env, 4-channel slab);
//   * by g++ as a serial emulation (wsmg_emul.cpp, tid0 = 0 / stride = 1) that the CPU
//     tests compare against the oracle.  The emulation is test infrastructure; nothing
//     in the product path calls it.
//
// Shared-memory plan (E=100, G=240: 231.6 KB of the 227 KB... see fused_smem_bytes):
//   X    [E*E] F4     rotated ego grid R, later the back-translated crop B
//   P    planar u32 keys [4][fan_cells] during the scatter, then F4[fan_cells];
//        after the first rotation the region is reused for the F ring + translate tables
//   Gst  2 x [BAND*WW] F4   cp.async-staged bands of the caller's global map window
//   tail baseE[E], rowoff[fan_rows], flags
#pragma once
#include "wsmg_math.h"

#if defined(__CUDACC__)
#define WSMG_BODY __device__ __forceinline__
#define WSMG_SYNC() __syncthreads()
#else
#define WSMG_BODY inline
#define WSMG_SYNC() ((void)0)
#endif

namespace wsmg {

constexpr int SLAB = 4;     // channels per CTA
constexpr int BAND = 8;     // window rows per fuse band
constexpr int RING = BAND + 2;

struct alignas(16) F4 { float v[4]; };
WSMG_HD F4 f4_zero() { F4 r; r.v[0] = r.v[1] = r.v[2] = r.v[3] = 0.0f; return r; }

struct FusedParams {
  const float* feat;        // [bs,C,Hf,Wf]
  const uint16_t* codes;    // [bs,Hf*Wf] packed fan codes from k_cells
  const float* gps;         // [bs,2]
  const float* compass;     // [bs,1]
  const float* trig;        // optional [bs,4]
  float* gmap;              // [n_maps,G,G,C]
  float* ego;               // [bs,C,E,E]
  float* proj_out;          // optional dump of the pre-rotation grid [bs,C,E,E]
  const float* proj_in;     // optional: take the grid from here instead of scattering
  int stop_after_scatter;   // stage API: return after writing proj_out
  int bs;
  Geo g;
};

struct SmemPlan {
  int x_off, p_off, p_bytes, gst_off, base_off, rowoff_off, flag_off, total;
  int npp;       // padded cells per key plane
  // views inside the P region after the first rotation
  int ring_off, tab_off;
};

WSMG_HD int align16(int x) { return (x + 15) & ~15; }

WSMG_HD SmemPlan make_plan(const Geo& g) {
  SmemPlan s;
  const int WW = g.E + 2;
  s.npp = (g.fan_cells + 3) & ~3;
  s.x_off = 0;
  s.p_off = align16(g.E * g.E * 16);
  int p_scatter = s.npp * 16;
  int ring_bytes = RING * WW * 16;
  int tab_bytes = align16((4 * WW + 4 * g.E) * 4);
  s.ring_off = s.p_off;
  s.tab_off = s.p_off + ring_bytes;
  s.p_bytes = p_scatter > ring_bytes + tab_bytes ? p_scatter : ring_bytes + tab_bytes;
  s.gst_off = s.p_off + align16(s.p_bytes);
  s.base_off = s.gst_off + 2 * BAND * WW * 16;
  s.rowoff_off = s.base_off + align16(g.E * 4);
  s.flag_off = s.rowoff_off + align16((g.fan_rows + 1) * 4);
  s.total = s.flag_off + 16;
  return s;
}

// ------------------------------------------------------------------ platform shims
#if defined(__CUDACC__)
__device__ __forceinline__ void smem_max(uint32_t* a, uint32_t v) { atomicMax(a, v); }
__device__ __forceinline__ F4 ld_stream4(const float* p) {
  float4 t = __ldcs(reinterpret_cast<const float4*>(p));
  F4 r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r;
}
__device__ __forceinline__ void st_stream(float* p, float v) { __stcs(p, v); }
__device__ __forceinline__ void async_copy16(void* dst_smem, const void* src, bool pred) {
  unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  int n = pred ? 16 : 0;   // src-size 0 => 16 bytes of zero fill, src not read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void async_copy4(void* dst_smem, const void* src, bool pred) {
  unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  int n = pred ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
#else
inline void smem_max(uint32_t* a, uint32_t v) { if (v > *a) *a = v; }
inline F4 ld_stream4(const float* p) { F4 r; for (int i = 0; i < 4; ++i) r.v[i] = p[i]; return r; }
inline void st_stream(float* p, float v) { *p = v; }
inline void async_copy16(void* dst, const void* src, bool pred) {
  if (pred) __builtin_memcpy(dst, src, 16); else __builtin_memset(dst, 0, 16);
}
inline void async_copy4(void* dst, const void* src, bool pred) {
  if (pred) __builtin_memcpy(dst, src, 4); else __builtin_memset(dst, 0, 4);
}
inline void async_commit() {}
template <int N> inline void async_wait() {}
#endif

// ------------------------------------------------------------------ the CTA body
// VEC: C % 4 == 0, so every (cell, slab) of the NHWC map is one aligned 16-byte word.
template <bool VEC>
WSMG_BODY void fused_body(const FusedParams& p, int block, unsigned char* smem, int tid0, int ts) {
  const Geo& g = p.g;
  const int E = g.E, G = g.G, C = g.C, WW = E + 2;
  const int HW = g.Hf * g.Wf;
  const int slabs = (C + SLAB - 1) / SLAB;
  const int b = block / slabs;
  const int c0 = (block - b * slabs) * SLAB;
  const int nch = (C - c0) < SLAB ? (C - c0) : SLAB;
  const SmemPlan sp = make_plan(g);

  F4* X = reinterpret_cast<F4*>(smem + sp.x_off);
  uint32_t* Pk = reinterpret_cast<uint32_t*>(smem + sp.p_off);
  F4* Pf = reinterpret_cast<F4*>(smem + sp.p_off);
  F4* Gst = reinterpret_cast<F4*>(smem + sp.gst_off);
  float* baseE = reinterpret_cast<float*>(smem + sp.base_off);
  int* rowoff = reinterpret_cast<int*>(smem + sp.rowoff_off);
  int* flags = reinterpret_cast<int*>(smem + sp.flag_off);

  // ---- pose scalars (every thread, redundantly) ------------------- rgb_mapping.py:34,45-51,57-63
  float gxc, gyc;
  gps_cell(g, p.gps[2 * b], p.gps[2 * b + 1], &gxc, &gyc);
  float sy = gxc - g.gcenter, sx = gyc - g.gcenter;
  const float qx = sx / g.gcenter, qy = sy / g.gcenter;     // retrieval pose; forward pose is the negation
  const float lim = (float)(4 * G);
  sy = fminf(fmaxf(sy, -lim), lim);                         // far-out (or NaN) poses: window leaves the map
  sx = fminf(fmaxf(sx, -lim), lim);
  const int u0 = (int)sy + g.paste_lo - 1;                  // global row / col of window cell (0,0)
  const int v0 = (int)sx + g.paste_lo - 1;
  float* gmap_b = p.gmap + (size_t)b * G * G * C;

  auto prefetch_band = [&](int k) {
    F4* dst = Gst + (k & 1) * BAND * WW;
    for (int t = tid0; t < BAND * WW; t += ts) {
      int rr = t / WW, vv = t - rr * WW, uu = k * BAND + rr;
      int u = u0 + uu, v = v0 + vv;
      bool inside = uu < WW && u >= 0 && u < G && v >= 0 && v < G;
      const float* src = gmap_b + ((size_t)(inside ? u : 0) * G + (inside ? v : 0)) * C + c0;
      if (VEC) {
        async_copy16(dst + t, src, inside);
      } else {
        for (int ch = 0; ch < SLAB; ++ch) async_copy4(&dst[t].v[ch], src + (ch < nch ? ch : 0), inside && ch < nch);
      }
    }
    async_commit();
  };
  const int NB = (WW + BAND - 1) / BAND;
  if (!p.stop_after_scatter) {        // the map window streams in underneath the scatter
    prefetch_band(0);
    if (NB > 1) prefetch_band(1);
  }

  // ---- tables ------------------------------------------------------------------------------
  for (int t = tid0; t < E; t += ts) baseE[t] = base_coord(t, E);
  for (int t = tid0; t <= g.fan_rows; t += ts) {
    int off = 0;
    for (int y = 0; y < t; ++y) off += fan_row_width(y, E);
    rowoff[t] = off;
  }
  if (tid0 == 0) flags[0] = 0;
  for (int t = tid0; t < SLAB * sp.npp; t += ts) Pk[t] = 0u;
  WSMG_SYNC();

  // ---- phase 1: scatter-max into the packed fan (rgb_mapping.py:210-225) -----------------------
  if (p.proj_in == nullptr) {
    const uint2* codes4 = reinterpret_cast<const uint2*>(p.codes + (size_t)b * HW);
    const float* plane0 = p.feat + ((size_t)b * C + c0) * HW;
    const int n4 = HW / 4;
    int saw_invalid = 0;
    for (int t = tid0; t < n4; t += 2 * ts) {
      // two 4-pixel groups per trip so that 8 feature loads are in flight per thread
      uint2 cc[2];
      bool live[2];
      F4 f[2][SLAB];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        int tt = t + h * ts;
        live[h] = tt < n4;
        cc[h].x = cc[h].y = 0xFFFFFFFFu;
        if (live[h]) {
#if defined(__CUDACC__)
          cc[h] = __ldg(codes4 + tt);
#else
          cc[h] = codes4[tt];
#endif
        }
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        int tt = t + h * ts;
        // packed codes of valid pixels are < CODE_OUTLIER; skip the feature read when none of the four writes
        bool any = live[h] && (((cc[h].x & 0xFFFFu) < CODE_OUTLIER) || ((cc[h].x >> 16) < CODE_OUTLIER) ||
                               ((cc[h].y & 0xFFFFu) < CODE_OUTLIER) || ((cc[h].y >> 16) < CODE_OUTLIER));
        live[h] = any;
#pragma unroll
        for (int ch = 0; ch < SLAB; ++ch)
          if (any && ch < nch) f[h][ch] = ld_stream4(plane0 + (size_t)ch * HW + 4 * (size_t)tt);
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t code[5];
        code[0] = cc[h].x & 0xFFFFu; code[1] = cc[h].x >> 16;
        code[2] = cc[h].y & 0xFFFFu; code[3] = cc[h].y >> 16; code[4] = 0xFFFFFFFFu;
        if ((t + h * ts) < n4 &&
            (code[0] >= CODE_OUTLIER || code[1] >= CODE_OUTLIER || code[2] >= CODE_OUTLIER || code[3] >= CODE_OUTLIER))
          saw_invalid = 1;
        if (!live[h]) continue;
#pragma unroll
        for (int ch = 0; ch < SLAB; ++ch) {
          if (ch >= nch) break;
          uint32_t* plane = Pk + ch * sp.npp;
          uint32_t run = 0u;
#pragma unroll
          for (int px = 0; px < 4; ++px) {
            if (code[px] < CODE_OUTLIER) {
              uint32_t k = f2key(f[h][ch].v[px]);
              run = k > run ? k : run;
              if (code[px + 1] != code[px]) { smem_max(plane + code[px], run); run = 0u; }
            }
          }
        }
      }
    }
    if (saw_invalid) flags[0] = 1;      // same value from every writer
  }
  WSMG_SYNC();

  // ---- phase 1b: keys -> finished floats, planar -> F4 per cell (rgb_mapping.py:228-230) -----
  if (p.proj_in == nullptr) {
    const uint32_t sentinel_key = f2key(SENTINEL);
    const bool inv = flags[0] != 0;
    for (int t = tid0; t < g.fan_cells; t += ts) {
      F4 v;
#pragma unroll
      for (int ch = 0; ch < SLAB; ++ch) {
        uint32_t k = Pk[ch * sp.npp + t];
        if (t == 0 && inv && k < sentinel_key) k = sentinel_key;   // invalid pixels write -1e16 to cell 0 (:207-212)
        v.v[ch] = ch < nch ? finish_cell(k) : 0.0f;
      }
      X[t] = v;
    }
    WSMG_SYNC();
    for (int t = tid0; t < g.fan_cells; t += ts) Pf[t] = X[t];
  } else {
    // stage API: load the projection (zero outside the fan by construction)
    const float* src = p.proj_in + ((size_t)b * C + c0) * E * E;
    for (int y = 0; y < g.fan_rows; ++y) {
      int xs = fan_x_lo(y), w = fan_row_width(y, E);
      for (int t = tid0; t < w; t += ts) {
        F4 v;
#pragma unroll
        for (int ch = 0; ch < SLAB; ++ch) v.v[ch] = ch < nch ? src[(size_t)ch * E * E + y * E + xs + t] : 0.0f;
        Pf[rowoff[y] + t] = v;
      }
    }
  }
  WSMG_SYNC();

  auto fan_at = [&](int y, int x) -> F4 {
    if (y < 0 || y >= g.fan_rows) return f4_zero();
    int xs = fan_x_lo(y);
    if (x < xs || x > fan_x_hi(y, E)) return f4_zero();
    return Pf[rowoff[y] + x - xs];
  };

  if (p.proj_out != nullptr) {
    float* dst = p.proj_out + ((size_t)b * C + c0) * E * E;
    for (int t = tid0; t < E * E; t += ts) {
      int y = t / E, x = t - y * E;
      F4 v = fan_at(y, x);
#pragma unroll
      for (int ch = 0; ch < SLAB; ++ch)
        if (ch < nch) dst[(size_t)ch * E * E + t] = v.v[ch];
    }
  }
  if (p.stop_after_scatter) return;

  // ---- phase 2: R = rotate(P, -compass) into X (rgb_mapping.py:37 -> :267 -> :239-250) ---------
  float cs, sn;
  if (p.trig != nullptr) { cs = p.trig[4 * b + 0]; sn = p.trig[4 * b + 1]; }
  else { float h = -p.compass[b]; sn = sinf(h); cs = cosf(h); }
  for (int t = tid0; t < E * E; t += ts) {
    int i = t / E, j = t - i * E;
    float ix, iy;
    rot_coords(baseE[j], baseE[i], cs, sn, g.half_e, &ix, &iy);
    Tap1D tx = make_tap(ix), ty = make_tap(iy);
    Weights w = make_weights(tx.w1, ty.w1);
    F4 a = fan_at(ty.i0, tx.i0), bb = fan_at(ty.i0, tx.i0 + 1);
    F4 c = fan_at(ty.i0 + 1, tx.i0), d = fan_at(ty.i0 + 1, tx.i0 + 1);
    F4 r;
#pragma unroll
    for (int ch = 0; ch < SLAB; ++ch) r.v[ch] = blend4(a.v[ch], bb.v[ch], c.v[ch], d.v[ch], w.nw, w.ne, w.sw, w.se);
    X[t] = r;
  }
  WSMG_SYNC();

  // ---- phase 3 tables: the two (separable) translations (rgb_mapping.py:45-53, 57-65) ----------
  int* colX0 = reinterpret_cast<int*>(smem + sp.tab_off);
  float* colW = reinterpret_cast<float*>(colX0 + WW);
  int* rowY0 = reinterpret_cast<int*>(colW + WW);
  float* rowW = reinterpret_cast<float*>(rowY0 + WW);
  int* bX0 = reinterpret_cast<int*>(rowW + WW);
  float* bWx = reinterpret_cast<float*>(bX0 + E);
  int* bY0 = reinterpret_cast<int*>(bWx + E);
  float* bWy = reinterpret_cast<float*>(bY0 + E);
  F4* ring = reinterpret_cast<F4*>(smem + sp.ring_off);
  for (int t = tid0; t < WW; t += ts) {
    int v = v0 + t, u = u0 + t;
    if (v >= 0 && v < G) {        // canvas column sampled by global column v, relative to the pasted ego grid
      Tap1D tp = make_tap(unnormalize(base_coord(v, G) + (-qx), g.half_g));
      colX0[t] = tp.i0 - g.paste_lo; colW[t] = tp.w1;
    } else { colX0[t] = -4; colW[t] = 0.0f; }
    if (u >= 0 && u < G) {
      Tap1D tp = make_tap(unnormalize(base_coord(u, G) + (-qy), g.half_g));
      rowY0[t] = tp.i0 - g.paste_lo; rowW[t] = tp.w1;
    } else { rowY0[t] = -4; rowW[t] = 0.0f; }
  }
  for (int t = tid0; t < E; t += ts) {   // global column / row sampled by crop cell t, relative to the window
    Tap1D tp = make_tap(unnormalize(base_coord(t + g.paste_lo, G) + qx, g.half_g));
    bX0[t] = tp.i0 - v0; bWx[t] = tp.w1;
    Tap1D tq = make_tap(unnormalize(base_coord(t + g.paste_lo, G) + qy, g.half_g));
    bY0[t] = tq.i0 - u0; bWy[t] = tq.w1;
  }

  auto rot_at = [&](int y, int x) -> F4 {
    if (y < 0 || y >= E || x < 0 || x >= E) return f4_zero();
    return X[y * E + x];
  };
  auto ring_at = [&](int row, int col) -> F4 {
    if (row < 0 || row >= WW || col < 0 || col >= WW) return f4_zero();
    return ring[(row % RING) * WW + col];
  };

  // ---- phase 3: banded translate + max-fuse (:53-56) and translate back + crop (:64-69) -------
  int p_lo = 0;
  for (int k = 0; k < NB; ++k) {
    if (k + 1 < NB) async_wait<1>(); else async_wait<0>();
    WSMG_SYNC();
    const F4* gst = Gst + (k & 1) * BAND * WW;
    for (int t = tid0; t < BAND * WW; t += ts) {
      int rr = t / WW, vv = t - rr * WW, uu = k * BAND + rr;
      if (uu >= WW) continue;
      int u = u0 + uu, v = v0 + vv;
      F4 f = f4_zero();
      if (u >= 0 && u < G && v >= 0 && v < G) {
        int ry = rowY0[uu], rx = colX0[vv];
        Weights w = make_weights(colW[vv], rowW[uu]);
        F4 a = rot_at(ry, rx), bb = rot_at(ry, rx + 1), c = rot_at(ry + 1, rx), d = rot_at(ry + 1, rx + 1);
        F4 old = gst[t];
#pragma unroll
        for (int ch = 0; ch < SLAB; ++ch) {
          float tv = blend4(a.v[ch], bb.v[ch], c.v[ch], d.v[ch], w.nw, w.ne, w.sw, w.se);
          f.v[ch] = fmaxf(old.v[ch], tv);
        }
        float* dst = gmap_b + ((size_t)u * G + v) * C + c0;
        if (VEC) {
          *reinterpret_cast<F4*>(dst) = f;
        } else {
#pragma unroll
          for (int ch = 0; ch < SLAB; ++ch)
            if (ch < nch) dst[ch] = f.v[ch];
        }
      }
      ring[(uu % RING) * WW + vv] = f;
    }
    WSMG_SYNC();
    if (k + 2 < NB) prefetch_band(k + 2);
    // B rows whose three source rows of F (and hence all readers of R row p) are done
    int done = (k + 1) * BAND < WW ? (k + 1) * BAND : WW;
    int p_hi = done - 2 > p_lo ? done - 2 : p_lo;
    for (int t = tid0; t < (p_hi - p_lo) * E; t += ts) {
      int pr = p_lo + t / E, q = t % E;
      int fy = bY0[pr], fx = bX0[q];
      Weights w = make_weights(bWx[q], bWy[pr]);
      F4 a = ring_at(fy, fx), bb = ring_at(fy, fx + 1), c = ring_at(fy + 1, fx), d = ring_at(fy + 1, fx + 1);
      F4 r;
#pragma unroll
      for (int ch = 0; ch < SLAB; ++ch) r.v[ch] = blend4(a.v[ch], bb.v[ch], c.v[ch], d.v[ch], w.nw, w.ne, w.sw, w.se);
      X[pr * E + q] = r;
    }
    p_lo = p_hi;
  }
  WSMG_SYNC();

  // ---- phase 4: ego = rotate(B, +compass), NCHW out (rgb_mapping.py:70) ----------------------
  if (p.trig != nullptr) { cs = p.trig[4 * b + 2]; sn = p.trig[4 * b + 3]; }
  else { float h = p.compass[b]; sn = sinf(h); cs = cosf(h); }
  float* ego_b = p.ego + ((size_t)b * C + c0) * E * E;
  for (int t = tid0; t < E * E; t += ts) {
    int i = t / E, j = t - i * E;
    float ix, iy;
    rot_coords(baseE[j], baseE[i], cs, sn, g.half_e, &ix, &iy);
    Tap1D tx = make_tap(ix), ty = make_tap(iy);
    Weights w = make_weights(tx.w1, ty.w1);
    F4 a = rot_at(ty.i0, tx.i0), bb = rot_at(ty.i0, tx.i0 + 1);
    F4 c = rot_at(ty.i0 + 1, tx.i0), d = rot_at(ty.i0 + 1, tx.i0 + 1);
#pragma unroll
    for (int ch = 0; ch < SLAB; ++ch)
      if (ch < nch) st_stream(ego_b + (size_t)ch * E * E + t, blend4(a.v[ch], bb.v[ch], c.v[ch], d.v[ch], w.nw, w.ne, w.sw, w.se));
  }
}

}  // namespace wsmg
