// Body of the fused map-update CTA, written once and compiled twice:
//   * by nvcc as the device code of k_fused (wsmg.cu), one CTA per (env, 4-channel slab);
//   * by g++ as a serial emulation (wsmg_emul.cpp, NT = 1) that the CPU tests compare against
//     the oracle.  The emulation is test infrastructure; nothing in the product path calls it.
//
// Design rules that came out of the ncu captures under profiles/ (the kernel is bound per SM -- instruction
// issue, LSU wavefronts and the barriers between phases at the 32 warps one 229 KB CTA leaves an SM -- not by DRAM):
//   * CTA size, ego/global size and the feature-plane stride are template constants for the reference
//     shapes; single-trip loops collapse to an `if`, divisions to multiplies, plane offsets to immediates;
//   * every bilinear tap is `table entry + table entry -> LDS.128`, index <= 0 meaning "zero": out-of-range
//     taps, taps whose weight is exactly 0 (two thirds of the translate rows/columns) and taps outside the
//     fan are never read; the blend runs on the packed fp32 pipe (FMUL2 / FFMA2, two channels per instruction);
//   * the NHWC map window is loaded with TMA boxes (no LSU) and only the cells an observation raises are
//     stored back, straight from the fuse; the feature planes move with per-thread cp.async slots staged in
//     the not-yet-used X buffer;
//   * the scatter reduces runs of equal cells in registers and issues plain shared atomics on signed keys
//     for which a non-negative float is its own key (profiles/: match_any+redux aggregation is 39x slower);
//     its loop is statically strided and unrolled so that queues and slots are register names.
//
// Round 2 measured nine restructurings of this body and kept none of them (summaries under profiles/r02_*, DESIGN.md
// section 5): two 512-thread CTAs per SM (crop rows in an L2-resident slot, tiled output rotation: +24 % instructions,
// 3.91 -> 4.45 ms), per-env plan tables (taps / weights precomputed once per env: the extra 16-byte load per cell costs
// more than the arithmetic it replaces), the output rotation run beside the band loop by the warps that do not own a
// window cell (the two phases stay additive), L2 prefetch of the scatter's feature chunks, a separate scatter kernel at
// two CTAs per SM, several slabs per CTA (the slab loop alone costs 0.29 ms: ptxas allocates the scatter's 64
// registers differently), the ego rows leaving through shared memory + bulk copies (400 small bulk copies per CTA cost
// 1 ms per 1024 envs), the first rotation over a list of the tiles that can see the fan.  What was kept: the band loop runs on the 26 warps that own a cell.
//
// Shared memory (E=100, G=240: 229 KB of the 227 KiB a CTA may opt in to):
//   X     [1 + E*E] F4    zero cell + (during the scatter) the per-thread cp.async feature slots, then the
//                         rotated ego grid R; its rows are overwritten by the crop B behind the fuse front
//   Z     1 F4            zero cell shared by the fan and the F ring (sits right before R2)
//   R2    scatter: planar signed-int keys [4][npp], then F4[fan_cells];
//         afterwards: F ring, `rr` window rows of WWP cells (128-byte aligned rows).  The caller's map
//         window arrives as TMA boxes {4 ch, WWP cols, TMA_ROWS rows} copied straight into ring rows
//         (mbarrier complete_tx; cells outside the map arrive as zeros) and is max-fused IN PLACE; raised
//         cells go back with one 16-byte st.global each.  Band 0 lands beyond the key planes so that it
//         streams in underneath the scatter.  (C % 4 != 0: cp.async loads instead of TMA.)
//   T     colT[WW], rowT[WW], bXT[E], bYT[E] (I4 each): the two separable translations
//   tail  baseE[E], fanrow[E+1] (later rowE[E+2]), ext[E], mbarriers, per-CTA scalars, hitrow[E]
#pragma once
#include "wsmg_math.h"
#if defined(__CUDACC__)
#include <cuda_fp16.h>
#endif

#if defined(__CUDACC__)
#define WSMG_BODY __device__ __forceinline__
#define WSMG_SYNC() __syncthreads()
#else
#define WSMG_BODY inline
#define WSMG_SYNC() ((void)0)
#endif

namespace wsmg {

constexpr int SLAB = 4;            // channels per CTA
constexpr int FUSED_NT = 1024;     // threads of a k_fused CTA
constexpr int SCATTER_STAGES = 2;  // cp.async feature slots per thread (staged in X, which is idle during the scatter)
constexpr int BAND = 8;            // window rows per fuse band
constexpr int TMA_ROWS = 4;        // rows per TMA box: BAND / TMA_ROWS loads per band; ring rows and S0 are multiples of it,
                                   // so a box never wraps around the ring
constexpr int NEG = -(1 << 24);    // "tap out of range": any index sum containing it is negative

struct alignas(16) F4 { float v[4]; };
struct alignas(16) I4 { int a, b, c, d; };
struct alignas(8) I2 { int a, b; };
WSMG_HD F4 f4_zero() { F4 r; r.v[0] = r.v[1] = r.v[2] = r.v[3] = 0.0f; return r; }
WSMG_HD int imax0(int a) { return a > 0 ? a : 0; }

constexpr int fan_cells_of(int E) {
  int rows = E / 2 + 1 < E ? E / 2 + 1 : E, n = 0;
  for (int y = 0; y < rows; ++y) {
    int lo = y - 2 > 0 ? y - 2 : 0, hi = E - y + 1 < E - 1 ? E - y + 1 : E - 1;
    n += hi - lo + 1 > 0 ? hi - lo + 1 : 0;
  }
  return n;
}
constexpr int npp_of(int fan_cells) { return (fan_cells + 3) & ~3; }

struct SmemPlan {
  int x_off, z_off, r2_off, tab_off, base_off, fanrow_off, ext_off, bar_off, total;
  int npp;        // words per key plane during the scatter
  int rr;         // ring rows
  int s0;         // ring slot of window row 0 (first slot beyond the key planes)
  int wwp;        // ring row stride in cells (row bytes are a multiple of 128 for TMA)
  uint32_t m_rr;  // fd_magic(rr)
};
constexpr int MAX_BANDS = 18;

WSMG_HD int align16(int x) { return (x + 15) & ~15; }

constexpr int wwp_of(int E) { return (E + 2 + 7) & ~7; }
constexpr int round_rows(int r) { return (r + TMA_ROWS - 1) / TMA_ROWS * TMA_ROWS; }
constexpr int s0_of(int E) { return round_rows((npp_of(fan_cells_of(E)) + wwp_of(E) - 1) / wwp_of(E)); }
constexpr int rr_of(int E) { return round_rows(4 * BAND + 2 > s0_of(E) + BAND ? 4 * BAND + 2 : s0_of(E) + BAND); }

WSMG_HD SmemPlan make_plan(const Geo& g) {
  SmemPlan s;
  const int WW = g.E + 2;
  s.wwp = (WW + 7) & ~7;
  s.npp = npp_of(g.fan_cells);
  s.s0 = round_rows((s.npp + s.wwp - 1) / s.wwp);
  s.rr = round_rows(4 * BAND + 2 > s.s0 + BAND ? 4 * BAND + 2 : s.s0 + BAND);
  s.m_rr = fd_magic(s.rr);
  s.x_off = 0;
  int x_cells = 1 + g.E * g.E;                               // X also hosts the scatter's staging slots
  if (x_cells < 1 + SCATTER_STAGES * SLAB * FUSED_NT) x_cells = 1 + SCATTER_STAGES * SLAB * FUSED_NT;
  s.r2_off = (x_cells * 16 + 16 + 127) & ~127;               // 128-byte aligned: TMA destination rows
  s.z_off = s.r2_off - 16;
  s.tab_off = s.r2_off + s.rr * s.wwp * 16;
  s.base_off = s.tab_off + (2 * WW + 2 * g.E) * 16;
  s.fanrow_off = s.base_off + align16(g.E * 4);
  s.ext_off = s.fanrow_off + align16((g.E + 2) * 8);         // fanrow[E+1]; after the first rotation the same bytes hold rowE[E+2]
  s.bar_off = s.ext_off + align16(g.E * 8);                  // ext[E]
  s.total = s.bar_off + MAX_BANDS * 8 + 32 + align16(g.E * 4);   // + per-CTA scalars (rotation sines / cosines, env flags)
                                                                  // + per-row column bounds of the first rotation
  return s;
}

#if defined(__CUDACC__)
struct alignas(64) TensorMapBlob { unsigned char bytes[128]; };   // a CUtensorMap, filled by the host glue
#else
struct TensorMapBlob { unsigned char bytes[8]; };
#endif

struct FusedParams {
  TensorMapBlob tmap;       // [n_maps,G,G,C] fp32, box {4, WWP, TMA_ROWS, 1}; filled for the TMA builds
  const uint32_t* env_flags; // [bs] bit0: the env has at least one pixel that does not write (k_cells, stage entry points)
  const uint32_t* block_flags; // whole step: [bs][flag_words] one word per k_cells block of the env (null: env_flags holds the OR)
  uint32_t* env_flags_out;  // whole step: [bs] receives the OR of the env's block words (written by the env's first slab CTA)
  int flag_words;
  const float* feat;        // [bs,C,Hf,Wf]
  const uint16_t* codes;    // [bs,Hf*Wf] packed fan codes from k_cells
  const float* gps;         // [bs,2]
  const float* compass;     // [bs,1]
  const float* trig;        // optional [bs,4]
  float* gmap;              // [n_maps,G,G,C]
  float* ego;               // [bs,C,E,E]
  uint16_t* ego_half;       // optional [bs,C,E,E] IEEE binary16 copy of ego (rollout store)
  const int32_t* env_slots; // optional [bs]: map row of frame b (default b)
  const int32_t* row_bounds; // optional [bs,E]: rot_row_bounds per env and R row (k_reset); null = no bounds
  const float* env_trig;    // optional [bs,4]: {cos, sin}(-compass), {cos, sin}(+compass) per env (k_reset); null = evaluate here
  uint32_t* status;         // optional uint32[2], device-accessible (pinned host memory): see wsmg_opts.status
  int n_maps;               // rows of the caller's map tensor (an env slot outside [0, n_maps) skips its frame)
  float* proj_out;          // optional dump of the pre-rotation grid [bs,C,E,E]
  const float* proj_in;     // optional: take the grid from here instead of scattering
  int stop_after_scatter;   // stage API: return after writing proj_out
  int debug_skip;           // read only by the -DWSMG_PHASE_SKIP profiling build (scripts/phase_split.py), see WSMG_SKIP
  int bs;
  Geo g;
  SmemPlan sp;
};

// ------------------------------------------------------------------ platform shims
#if defined(__CUDACC__)
__device__ __forceinline__ void smem_max(int32_t* a, int32_t v) { atomicMax(a, v); }
// true when no active lane of the (converged) warp has the sign bit set in `bits`
__device__ __forceinline__ bool warp_all_nonneg(int32_t bits) { return __ballot_sync(__activemask(), bits < 0) == 0u; }
__device__ __forceinline__ uint2 ld_codes(const uint2* p) { return __ldg(p); }
__device__ __forceinline__ void st_stream(float* p, float v) { __stcs(p, v); }
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
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
__device__ __forceinline__ uint16_t f2half_bits(float f) { return __half_as_ushort(__float2half_rn(f)); }
__device__ __forceinline__ unsigned saddr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// -- mbarrier / TMA (cp.async.bulk.tensor) -------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(saddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(saddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  unsigned done = 0;
  while (!done) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(saddr(bar)), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void tma_load_box(void* dst, const void* tmap, int c, int v, int u, int b, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n"
               ::"r"(saddr(dst)), "l"(tmap), "r"(c), "r"(v), "r"(u), "r"(b), "r"(saddr(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void group_sync(int id, int threads) { asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(threads) : "memory"); }
#else
inline void smem_max(int32_t* a, int32_t v) { if (v > *a) *a = v; }
inline bool warp_all_nonneg(int32_t bits) { return bits >= 0; }
inline uint2 ld_codes(const uint2* p) { return *p; }
inline void st_stream(float* p, float v) { *p = v; }
inline void grid_dependency_wait() {}
inline void async_copy16(void* dst, const void* src, bool pred) {
  if (pred) __builtin_memcpy(dst, src, 16); else __builtin_memset(dst, 0, 16);
}
inline void async_copy4(void* dst, const void* src, bool pred) {
  if (pred) __builtin_memcpy(dst, src, 4); else __builtin_memset(dst, 0, 4);
}
inline void async_commit() {}
template <int N> inline void async_wait() {}
inline uint16_t f2half_bits(float f) {        // round-to-nearest-even binary16, like numpy's astype(float16)
  _Float16 h = (_Float16)f; uint16_t u; __builtin_memcpy(&u, &h, 2); return u;
}
inline void mbar_init(uint64_t*, int) {}
inline void mbar_init_fence() {}
inline void mbar_expect_tx(uint64_t*, unsigned) {}
inline void mbar_wait(uint64_t*, unsigned) {}
inline void tma_load_box(void*, const void*, int, int, int, int, uint64_t*) {}
inline void fence_proxy_async() {}
inline void group_sync(int, int) {}
#endif

// Tap that skips the shared-memory read when the index says "zero cell" (out of range, or a tap whose
// bilinear weight is exactly 0 -- its product is an exact 0, so dropping the load is bit-exact for finite data).
WSMG_HD F4 tap(const F4* base, int idx) { return idx > 0 ? base[idx] : f4_zero(); }

// Lane -> cell mapping of the two rotations.  A warp covers an 8-column x 4-row tile of the E x E grid and
// each quarter-warp (the unit LDS.128 is issued in) a 4 x 2 block of it: the four bilinear taps of a block
// land on ~1.6 distinct 16-byte bank groups per wavefront instead of ~2.1 for 8 cells in a row (simulated
// over random headings and confirmed by ncu), and a row of the tile is still one full 32-byte sector of
// the NCHW output.  slot = tile * 32 + lane; returns false for the padding of ragged tiles.
WSMG_HD bool tile_cell(int slot, int E, int* i, int* j, uint32_t m_tiles = 0u) {
  const int tile = slot >> 5, l = slot & 31;
  const int tiles_x = (E + 7) >> 3;
  const int band = fd_div(tile, tiles_x, m_tiles), tcol = tile - band * tiles_x;
  const int q = l >> 3, k = l & 7;
  *j = 8 * tcol + 4 * (q & 1) + (k & 3);
  *i = 4 * band + 2 * (q >> 1) + (k >> 2);
  return *j < E && *i < E;
}
WSMG_HD int tile_slots(int E) { return ((E + 7) >> 3) * ((E + 3) >> 2) * 32; }

#if defined(__CUDA_ARCH__)
// Two channels per instruction: Blackwell's packed-fp32 pipe (FMUL2 / FFMA2) evaluates the blend4 chain of
// wsmg_math.h -- r = a*nw; r = fma(b, ne, r); r = fma(c, sw, r); r = fma(d, se, r), each step rounded to nearest --
// on a register pair, bit-identical per channel, at half the issue slots (the kernel is issue-bound).
__device__ __forceinline__ void blend_pair(float a0, float a1, float b0, float b1, float c0, float c1, float d0, float d1,
                                           const Weights& w, float* r0, float* r1) {
  asm("{\n .reg .b64 a, b, c, d, wn, we, ws, wd, r;\n"
      " mov.b64 a, {%2, %3};\n mov.b64 b, {%4, %5};\n mov.b64 c, {%6, %7};\n mov.b64 d, {%8, %9};\n"
      " mov.b64 wn, {%10, %10};\n mov.b64 we, {%11, %11};\n mov.b64 ws, {%12, %12};\n mov.b64 wd, {%13, %13};\n"
      " mul.rn.f32x2 r, a, wn;\n fma.rn.f32x2 r, b, we, r;\n fma.rn.f32x2 r, c, ws, r;\n fma.rn.f32x2 r, d, wd, r;\n"
      " mov.b64 {%0, %1}, r;\n}\n"
      : "=f"(*r0), "=f"(*r1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1), "f"(d0), "f"(d1), "f"(w.nw), "f"(w.ne), "f"(w.sw), "f"(w.se));
}
#endif

WSMG_HD F4 blend_f4(const F4& a, const F4& b, const F4& c, const F4& d, const Weights& w) {
  F4 r;
#if defined(__CUDA_ARCH__)
  static_assert(SLAB == 4, "blend_f4 pairs the four channels of a slab");
  blend_pair(a.v[0], a.v[1], b.v[0], b.v[1], c.v[0], c.v[1], d.v[0], d.v[1], w, &r.v[0], &r.v[1]);
  blend_pair(a.v[2], a.v[3], b.v[2], b.v[3], c.v[2], c.v[3], d.v[2], d.v[3], w, &r.v[2], &r.v[3]);
#else
#pragma unroll
  for (int ch = 0; ch < SLAB; ++ch) r.v[ch] = blend4(a.v[ch], b.v[ch], c.v[ch], d.v[ch], w.nw, w.ne, w.sw, w.se);
#endif
  return r;
}

// Column bounds of the first rotation, one word per R row: first | (last + 1) << 16.
// R(i,j) samples the fan at (ix, iy) = (h + u*cs + v*sn, h - u*sn + v*cs), u = j - h, v = i - h, h = (E-1)/2 (the
// affine_grid / unnormalize chain in real arithmetic).  One of its four taps can lie in the fan only if
// -1 <= iy < fan_rows, ix - iy >= -4 and ix + iy < E + 3; per row each condition is a bound on u.  With 1.5 cells
// of slack (float rounding is ~1e-4 cells) this is a superset of the cells that hit, so skipping the rest
// changes nothing: they are the exact zeros the full evaluation would store.  ~70 % of the grid.
// Evaluated once per env by k_reset (not per slab CTA), read by k_fused.
WSMG_HD int32_t rot_row_bounds(const Geo& g, float cs1, float sn1, int row) {
  const int E = g.E;
  const float h = g.half, m = 1.5f, v = (float)row - h;
  float lo = -4.0f * (float)E, hi = 4.0f * (float)E;
  bool none = false;
  auto bound = [&](float a, float b) {                        // a*u >= b
    if (a > 1e-3f) lo = fmaxf(lo, b / a);
    else if (a < -1e-3f) hi = fminf(hi, b / a);
    else if (b > 1.0f) none = true;                          // |a*u| <= 0.1 everywhere on the grid
  };
  bound(-sn1, -1.0f - m - h - v * cs1);                                      // iy >= -1 - m
  bound(sn1, -((float)g.fan_rows + m) + h + v * cs1);                        // iy <= fan_rows + m
  bound(cs1 + sn1, -4.0f - m - v * (sn1 - cs1));                             // ix - iy >= -4 - m
  bound(sn1 - cs1, -((float)E + 3.0f + m) + 2.0f * h + v * (sn1 + cs1));     // ix + iy <= E + 3 + m
  int jlo = (int)floorf(lo + h) - 1, jhi1 = (int)ceilf(hi + h) + 2;
  jlo = jlo < 0 ? 0 : jlo; jhi1 = jhi1 > E ? E : jhi1;
  if (none || jlo >= jhi1) { jlo = 0; jhi1 = 0; }
  return jlo | (jhi1 << 16);
}

// ------------------------------------------------------------------ the CTA body
// NT: threads per CTA (1 in the emulation).  CE/CG/CHW > 0: ego size, global size and Hf*Wf known
// at compile time (the reference's 100 / 240 / 224*224); 0: read from p.g.
// VEC: C % 4 == 0, so every (cell, slab) of the NHWC map is one aligned 16-byte word.
// TMA: the map window moves through cp.async.bulk.tensor (needs VEC); else cp.async + st.global.
// FEAT: how the features arrive.  FEAT_NCHW: [bs,C,Hf,Wf]; FEAT_POOL: NCHW with C_in != C, the channel pool of
// rgb_mapping.py:81-84 runs inside the scatter; FEAT_NHWC: [bs,Hf,Wf,C] (a channels_last producer) -- a
// pixel's four slab channels are one 16-byte word, so the staging slots hold one F4 per PIXEL instead of per channel.
// Phase-skipping switch of the profiling build (build.py --phase-skip -> lib/libwsmg_phaseskip.so): bit 1 scatter,
// 2 first rotation, 4 band loop, 8 output rotation, 16 crop, 32 fuse, 64 TMA-arrival wait, 512 key decode,
// 1024 translation tables, 2048 key-plane init, 4096 return at entry (launch cost of an empty CTA),
// 16384 the scatter's shared-memory atomics, 32768 the scatter's feature copies.  Results are
// garbage with any bit set; only the timing means something.  The product build compiles it away.
#if defined(WSMG_PHASE_SKIP)
#define WSMG_SKIP(bit) ((p.debug_skip & (bit)) != 0)
#else
#define WSMG_SKIP(bit) false
#endif

constexpr int FEAT_NCHW = 0, FEAT_POOL = 1, FEAT_NHWC = 2;

template <int NT, int CE, int CG, int CHW, bool VEC, bool TMA_BUILD, int FEAT>
WSMG_BODY void fused_body(const FusedParams& p, int block, unsigned char* smem, const int tid) {
  const Geo& g = p.g;
  const SmemPlan& sp = p.sp;
  const int E = CE > 0 ? CE : g.E;
  const int G = CG > 0 ? CG : g.G;
  const int HW = CHW > 0 ? CHW : g.Hf * g.Wf;
  const int C = g.C, WW = E + 2, EE = E * E;
  const int fan_cells = CE > 0 ? fan_cells_of(CE) : g.fan_cells;
  const int npp = CE > 0 ? npp_of(fan_cells_of(CE)) : sp.npp;
  const int WWP = CE > 0 ? wwp_of(CE) : sp.wwp;
  const int RR = CE > 0 ? rr_of(CE) : sp.rr;
  const int S0 = CE > 0 ? s0_of(CE) : sp.s0;
  const int slabs = (C + SLAB - 1) / SLAB;
  const int b = block / slabs;
  const int c0 = (block - b * slabs) * SLAB;
  const int nch = VEC ? SLAB : ((C - c0) < SLAB ? (C - c0) : SLAB);
  const int paste_lo = G / 2 - E / 2;
  const float half_e = (float)E / 2.0f, half_g = (float)G / 2.0f, gcenter = (float)(G / 2);
  const int NB = (WW + BAND - 1) / BAND;          // <= MAX_BANDS - 1 (validated on the host)
  // divisions by run-time geometry go through fd_div (compile-time geometry: plain constants, m = 0)
  const uint32_t mE = CE > 0 ? 0u : g.m_E, mWW = CE > 0 ? 0u : g.m_WW, mT = CE > 0 ? 0u : g.m_tiles, mRR = CE > 0 ? 0u : sp.m_rr;
  auto ring_slot = [&](int r) -> int { const int x = r + S0; return x - fd_div(x, RR, mRR) * RR; };   // (r + S0) % RR

  F4* X = reinterpret_cast<F4*>(smem + sp.x_off);            // X[0] = zero cell, R/B(y,x) at X[1 + y*E + x]
  int32_t* Pk = reinterpret_cast<int32_t*>(smem + sp.r2_off);
  F4* Pf = reinterpret_cast<F4*>(smem + sp.z_off);           // Pf[0] = zero cell, fan cell c at Pf[1 + c]
  F4* ring = reinterpret_cast<F4*>(smem + sp.z_off);         // ring[0] = zero cell, (slot, col) at ring[1 + slot*WWP + col]
  I4* colT = reinterpret_cast<I4*>(smem + sp.tab_off);
  I4* rowT = colT + WW;
  I4* bXT = rowT + WW;
  I4* bYT = bXT + E;
  float* baseE = reinterpret_cast<float*>(smem + sp.base_off);
  I2* fanrow = reinterpret_cast<I2*>(smem + sp.fanrow_off);   // per grid row y (entry E = "no such row"): {1+rowoff-xs, xs | xe<<16}
  I2* ext = reinterpret_cast<I2*>(smem + sp.ext_off);         // per R row: [first, last] column with a tap inside the fan
  I2* rowE = fanrow;                                          // per window row: merged extent of its two source R rows (fanrow is dead by then)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sp.bar_off);   // one mbarrier per band (TMA)
  float* scal = reinterpret_cast<float*>(bars + MAX_BANDS);   // {cos, sin}(-compass), {cos, sin}(+compass), env flags
  int32_t* hitrow = reinterpret_cast<int32_t*>(scal + 8);      // per R row: first | (last + 1) << 16 column that can see the fan
  if (WSMG_SKIP(4096)) return;
  // Launched as a programmatic dependent of k_cells: everything below reads what that grid (and k_reset before it)
  // wrote -- codes, flags, rotation tables, the reset map.  A no-op for a plain launch.
  grid_dependency_wait();

  // ---- per-env scalars that later phases need: fetched and evaluated now by three lanes of three different
  // warps, parked in shared memory -- their global-memory latency would otherwise stall the whole CTA right
  // after a barrier (rgb_mapping.py:37,70 -> :239-245)
  {
    const int t_r1 = 0, t_r2 = NT > 64 ? 32 : 0, t_fl = NT > 64 ? 64 : 0;
    if (!p.stop_after_scatter && p.env_trig != nullptr) {          // evaluated once per env by k_reset
      if (tid < 4) scal[tid] = p.env_trig[4 * b + tid];
    } else if (!p.stop_after_scatter) {
      if (tid == t_r1) {
        float cs_, sn_;
        if (p.trig != nullptr) { cs_ = p.trig[4 * b + 0]; sn_ = p.trig[4 * b + 1]; }
        else { float h = -p.compass[b]; sn_ = sinf(h); cs_ = cosf(h); }
        scal[0] = cs_; scal[1] = sn_;
      }
      if (tid == t_r2) {
        float cs_, sn_;
        if (p.trig != nullptr) { cs_ = p.trig[4 * b + 2]; sn_ = p.trig[4 * b + 3]; }
        else { float h = p.compass[b]; sn_ = sinf(h); cs_ = cosf(h); }
        scal[2] = cs_; scal[3] = sn_;
      }
    }
    if (tid == t_fl && p.proj_in == nullptr) {
      uint32_t fl = 0u;
      if (p.block_flags != nullptr) {
        for (int w = 0; w < p.flag_words; ++w) fl |= p.block_flags[(size_t)b * p.flag_words + w];
        if (block % ((C + SLAB - 1) / SLAB) == 0) p.env_flags_out[b] = fl;      // per-env summary (include/wsmg.h: wsmg_scratch_flags_offset)
      } else {
        fl = p.env_flags[b];
      }
      scal[4] = as_float((int)fl);
    }
  }

  // ---- pose scalars (every thread, redundantly) ------------------- rgb_mapping.py:34,45-51,57-63
  float qx = 0.f, qy = 0.f;
  int u0 = 0, v0 = 0;
  float* gmap_b = nullptr;
  const int mrow = p.env_slots != nullptr ? p.env_slots[b] : b;      // row of the caller's map tensor
  if (!p.stop_after_scatter) {
    if ((unsigned)mrow >= (unsigned)p.n_maps) {               // not a row of the map: skip the frame (k_reset flags it)
      if (tid == 0 && p.status != nullptr) p.status[1] = 1u;
      return;
    }
    float gxc, gyc;
    gps_cell(g, p.gps[2 * b], p.gps[2 * b + 1], &gxc, &gyc);
    float sy = gxc - gcenter, sx = gyc - gcenter;
    qx = sx / gcenter; qy = sy / gcenter;                   // retrieval pose; the forward pose is its negation
    const float lim = (float)(4 * G);
    sy = fminf(fmaxf(sy, -lim), lim);                       // far-out (or NaN) poses: the window leaves the map
    sx = fminf(fmaxf(sx, -lim), lim);
    u0 = (int)sy + paste_lo - 1;                            // global row / col of window cell (0,0)
    v0 = (int)sx + paste_lo - 1;
    gmap_b = p.gmap + (size_t)mrow * G * G * C + c0;
  }
  // The window is only ever LOADED through TMA, and a tiled load zero-fills whatever lies outside the tensor
  // (negative coordinates included): windows clipped by the map border need no special path.
  constexpr bool TMA = TMA_BUILD;

  // Band k of the caller's map window -> its ring rows; cells outside the map arrive as zeros.
  // TMA: boxes of TMA_ROWS x WWP cells, issued by the first lanes of the last warp (it owns no window / crop
  // cell).  Columns beyond the window and rows beyond its last one are loaded (or zero-filled outside the
  // tensor) and never read.
  // The band loop (phase 3) gives every thread at most one window cell and one crop cell per trip: BAND * (E + 2) cells.
  // Only the warps that own one run it, paced by a named barrier of their own -- the rest would just lengthen every
  // trip's barrier (measured: 3.505 -> 3.452 ms per 1024-env step with 26 instead of 32 warps at E = 100).
  const int LT = NT >= 64 ? (((BAND * WW + 31) / 32 * 32) < NT ? ((BAND * WW + 31) / 32 * 32) : NT) : NT;   // loop threads
  const int tma_lane = tid - (LT - 32);                     // 0..31 in the TMA warp (the last loop warp), else outside [0, 32)
  auto prefetch_band = [&](int k) {
    if (TMA) {
      if ((unsigned)tma_lane < 32u) {
        const int left = WW - k * BAND;
        const int boxes = left >= BAND ? BAND / TMA_ROWS : (left + TMA_ROWS - 1) / TMA_ROWS;
        if (tma_lane == 0) mbar_expect_tx(&bars[k], (unsigned)(boxes * TMA_ROWS * WWP * 16));
#if defined(__CUDACC__)
        __syncwarp();
#endif
        if (tma_lane < boxes) {
          const int uu = k * BAND + tma_lane * TMA_ROWS;
          tma_load_box(ring + 1 + ring_slot(uu) * WWP, &p.tmap, c0, v0, u0 + uu, mrow, &bars[k]);
        }
      }
    } else if (VEC) {
      for (int t = tid; t < BAND * WW; t += NT) {
        int rr = fd_div(t, WW, mWW), vv = t - rr * WW, uu = k * BAND + rr;
        if (uu < WW) {
          int u = u0 + uu, v = v0 + vv;
          bool inside = (unsigned)u < (unsigned)G && (unsigned)v < (unsigned)G;
          const float* src = gmap_b + (inside ? ((size_t)u * G + v) * C : 0);
          async_copy16(ring + 1 + ring_slot(uu) * WWP + vv, src, inside);
        }
      }
      async_commit();
    } else {
      // C % 4 != 0: the cells are only 4-byte aligned, so the window moves one float per copy.  Consecutive lanes take the
      // consecutive floats of a cell (item = cell * 4 + channel): a warp instruction then touches 8 cells' 16-byte
      // pieces instead of one float in each of 32 cells -- a quarter of the L1 tag lookups.
      for (int t = tid; t < BAND * WW * SLAB; t += NT) {
        const int cell = t / SLAB, ch = t - cell * SLAB;
        int rr = fd_div(cell, WW, mWW), vv = cell - rr * WW, uu = k * BAND + rr;
        if (uu < WW) {
          int u = u0 + uu, v = v0 + vv;
          bool inside = (unsigned)u < (unsigned)G && (unsigned)v < (unsigned)G && ch < nch;
          const float* src = gmap_b + (inside ? ((size_t)u * G + v) * C + ch : 0);
          async_copy4(&(ring + 1 + ring_slot(uu) * WWP + vv)->v[ch], src, inside);
        }
      }
      async_commit();
    }
  };
  if (!p.stop_after_scatter) {
    if (TMA && (unsigned)tma_lane < 32u) {
      if (tma_lane == 0) {
        for (int k = 0; k < NB; ++k) mbar_init(&bars[k], 1);
        mbar_init_fence();
      }
#if defined(__CUDACC__)
      __syncwarp();
#endif
    }
    prefetch_band(0);              // streams in underneath the scatter (slots beyond the key planes)
  }

  // ---- tables ------------------------------------------------------------------------------
  for (int t = tid; t < E; t += NT) baseE[t] = base_coord(t, E);
  for (int t = tid; t <= E; t += NT) {
    I2 fr; fr.a = 0; fr.b = 1;                              // xs = 1 > xe = 0: empty row
    if (t < g.fan_rows) {
      int xs = fan_x_lo(t), xe = fan_x_hi(t, E);
      fr.a = 1 + fan_row_offset(t, E) - xs; fr.b = xs | (xe << 16);
    }
    fanrow[t] = fr;
  }
  if (tid == 0) { X[0] = f4_zero(); Pf[0] = f4_zero(); }
  for (int t = tid; t < E; t += NT) { I2 e; e.a = E; e.b = -1; ext[t] = e; }     // empty extent
  for (int t = tid; t < E; t += NT) hitrow[t] = p.row_bounds != nullptr ? p.row_bounds[(size_t)b * E + t] : (E << 16);
  for (int t = tid; t < (WSMG_SKIP(2048) ? 0 : SLAB * npp); t += NT) Pk[t] = KEY_EMPTY;
  // (the barrier that publishes these tables and the empty key planes sits inside the scatter, after its first
  // code fetches and feature copies have been issued: their L2 / DRAM latency overlaps the prologue)

  // ---- phase 1: scatter-max into the packed fan (rgb_mapping.py:210-225) -----------------------
  const bool do_scatter = p.proj_in == nullptr && !WSMG_SKIP(1);
  if (!do_scatter) WSMG_SYNC();
  if (do_scatter) {
    const uint2* codes4 = reinterpret_cast<const uint2*>(p.codes + (size_t)b * HW);
    // Channel pool (rgb_mapping.py:81-84) fused: when Cin != C the scatter runs once per input plane of a bin,
    // all passes reducing into the same key plane (max over channels commutes with the max-scatter).
    const int Cin = g.Cin;
    constexpr bool pool = FEAT == FEAT_POOL;
    constexpr bool nhwc = FEAT == FEAT_NHWC;
    static_assert(!nhwc || VEC, "NHWC features need C % 4 == 0");
    int bin_lo[SLAB], bin_n[SLAB], passes = 1;
#pragma unroll
    for (int ch = 0; ch < SLAB; ++ch) {
      const int k = c0 + (ch < nch ? ch : 0);
      bin_lo[ch] = pool ? pool_start(k, Cin, C) : k;
      bin_n[ch] = pool ? pool_end(k, Cin, C) - bin_lo[ch] : 1;
      passes = bin_n[ch] > passes ? bin_n[ch] : passes;
    }
    const float* feat_b = p.feat + (size_t)b * Cin * HW;     // (the same element count per env in either layout)
    const int n4 = HW / 4;
    for (int pass = 0; pass < passes; ++pass) {
    size_t plane_off[SLAB];                                   // element offset of this pass's input plane per output channel
#pragma unroll
    for (int ch = 0; ch < SLAB; ++ch) plane_off[ch] = (size_t)(bin_lo[ch] + (pass < bin_n[ch] ? pass : bin_n[ch] - 1)) * HW;
    // Feature staging through shared memory: X is idle until phase 2, so every thread owns STAGES private
    // 64-byte slots in it ([slot][channel][thread] float4s: conflict-free) and keeps STAGES groups (4 pixels x 4
    // channels) in flight with cp.async while it reduces the oldest one -- no registers are tied up by loads in
    // flight and no barrier is involved (a thread only ever reads what it copied itself).
    // Codes run STAGES steps ahead of the features, so their L2 latency never sits in front of a feature copy.
    // Valid pixels carry a fan cell (< CODE_OUTLIER): every u16 lane of (x & y) has its top 15 bits set iff none of
    // the four pixels writes -- such a group skips its feature read.
    constexpr int STAGES = SCATTER_STAGES;
    static_assert(NT <= FUSED_NT, "staging slots are sized for FUSED_NT threads");
    F4* stage = X + 1;                                        // [STAGES][SLAB][NT]
    constexpr int QD = STAGES + 3;                            // queue depth: the codes run 3 steps ahead of the feature copies,
    uint2 cq[QD];                                             // so even the short iterations of the dead zone hide their L2 latency
    auto fetch_codes = [&](int tt) -> uint2 {
      uint2 c; c.x = c.y = 0xFFFFFFFFu;
      if (tt < n4) c = ld_codes(codes4 + tt);
      return c;
    };
    auto is_live = [](uint2 c) { return ((c.x & c.y) | 0x00010001u) != 0xFFFFFFFFu; };   // (codes past the end are 0xFFFF)
    const float* feat_slab = feat_b + (size_t)c0 * HW;        // plane of the slab's first channel (no pool)
    auto issue = [&](int slot, int tt, uint2 c) {
      if (is_live(c)) {
        F4* dst = stage + slot * SLAB * NT + tid;
        if (nhwc) {                           // pixel 4*tt + px, channels c0..c0+3: one aligned 16-byte word each
          const float* src = feat_b + 4 * (size_t)tt * Cin + c0;
#pragma unroll
          for (int px = 0; px < 4; ++px)
            if (!WSMG_SKIP(32768)) async_copy16(dst + px * NT, src + (size_t)px * Cin, true);
        } else {
          const float* src = (pool ? feat_b : feat_slab) + 4 * (size_t)tt;
#pragma unroll
          for (int ch = 0; ch < SLAB; ++ch)   // without the pool the four planes are compile-time offsets of one pointer
            if (ch < nch && !WSMG_SKIP(32768)) async_copy16(dst + ch * NT, src + (pool ? plane_off[ch] : (size_t)ch * HW), true);
        }
      }
      async_commit();                                          // one group per step, live or not: wait counts stay exact
    };
    // Work distribution: warp w takes the chunks of 32 consecutive groups w, w + warps, w + 2*warps, ... (a thread's
    // group in chunk c is c*32 + lane), which samples every image row band evenly.  (Pulling chunks from a shared counter was measured at 1, 2, 4 and 8 chunks per pull:
    // 1.5-4.5 % slower than this static striding.)
    // With NT a multiple of 32 that is simply: the thread's k-th group is tid + k*NT.
    int t_head = tid;                                          // group index of cq[0]; cq[s] belongs to t_head + s*NT
#pragma unroll
    for (int s_ = 0; s_ < QD; ++s_) cq[s_] = fetch_codes(t_head + s_ * NT);
#pragma unroll
    for (int s_ = 0; s_ < STAGES; ++s_) issue(s_, t_head + s_ * NT, cq[s_]);
    if (pass == 0) WSMG_SYNC();                                // key planes are empty: the first atomic may go
    // unrolled by lcm(QD, STAGES): the code queue's rotation and the slot index turn into register renaming and
    // immediates (measured: -3 % kernel time over the rolled loop; unroll 2 or a shorter / longer queue are slower)
    constexpr int UNR = QD * STAGES;
    static_assert(QD % STAGES != 0, "unroll factor assumes coprime queue depth and stage count");
#pragma unroll UNR
    for (int it = 0; t_head - (tid % 32) < n4; ++it) {        // warp-uniform: the warp's first lane still has a group
      const int slot = it % STAGES;
      const uint2 cc = cq[0];
      async_wait<STAGES - 1>();                                // this group's copies have landed
      if (is_live(cc)) {
        F4 f[SLAB];
#pragma unroll
        for (int ch = 0; ch < SLAB; ++ch) f[ch] = ch < nch ? stage[(slot * SLAB + ch) * NT + tid] : f4_zero();
        if (nhwc) {                           // slots hold pixels: transpose to f[channel].v[pixel] (register renaming)
          F4 gq[4];
#pragma unroll
          for (int px = 0; px < 4; ++px) gq[px] = f[px];
#pragma unroll
          for (int ch = 0; ch < SLAB; ++ch)
#pragma unroll
            for (int px = 0; px < 4; ++px) f[ch].v[px] = gq[px].v[ch];
        }
        // Runs of equal codes are reduced in registers with a predicated running max (the sign of a zero is
        // irrelevant, finish_cell() turns -0 into +0 like the reference): one shared-memory reduction per run and
        // channel.  A pixel that does not write never flushes, and the run restarts whenever the code changes,
        // so its value needs no masking.  When all 16 values are non-negative (post-ReLU features) their bit
        // patterns already are the keys.
        unsigned code[5];
        code[0] = cc.x & 0xFFFFu; code[1] = cc.x >> 16;
        code[2] = cc.y & 0xFFFFu; code[3] = cc.y >> 16; code[4] = 0xFFFFFFFFu;
        int32_t sign = 0;
#pragma unroll
        for (int ch = 0; ch < SLAB; ++ch)
          if (ch < nch) sign |= f_bits(f[ch].v[0]) | f_bits(f[ch].v[1]) | f_bits(f[ch].v[2]) | f_bits(f[ch].v[3]);
        // warp-uniform, so the common case (every lane non-negative) carries no key conversion at all
        const bool nonneg = warp_all_nonneg(sign);
        auto reduce_group = [&](auto to_key) {
#pragma unroll
          for (int px = 0; px < 4; ++px) {
            if (px > 0 && code[px] == code[px - 1]) {           // running max of the run, in place (predicated FMNMX)
#pragma unroll
              for (int ch = 0; ch < SLAB; ++ch) f[ch].v[px] = fmaxf(f[ch].v[px - 1], f[ch].v[px]);
            }
            // ptxas turns a predicated shared atomic into its own branch region, so one branch per pixel (the
            // predicate is shared by the four channel planes) is the cheapest form
            if (code[px] < CODE_OUTLIER && code[px] != code[px + 1]) {
              int32_t* cell = Pk + code[px];
#pragma unroll
              for (int ch = 0; ch < SLAB; ++ch)
                if (ch < nch && !WSMG_SKIP(16384)) smem_max(cell + ch * npp, to_key(f[ch].v[px]));
            }
          }
        };
        if (nonneg) reduce_group([](float v) { return f_bits(v); });
        else reduce_group([](float v) { return f2key(v); });
      }
      // refill the slot just consumed with the group STAGES steps ahead; advance the queues.  (Refilling before
      // the reduction, with the values already in registers, was measured: 2.5 % slower.)
#pragma unroll
      for (int s_ = 0; s_ + 1 < QD; ++s_) cq[s_] = cq[s_ + 1];
      t_head += NT;
      issue(slot, t_head + (STAGES - 1) * NT, cq[STAGES - 1]);
      cq[QD - 1] = fetch_codes(t_head + (QD - 1) * NT);
    }
    async_wait<0>();
    }   // pass
  }
  WSMG_SYNC();

  // ---- phase 1b: keys -> finished floats, planar -> F4 per cell (rgb_mapping.py:228-230) -----
  if (p.proj_in == nullptr) {
    const int32_t sentinel_key = f2key(SENTINEL);
    const bool inv = (as_int(scal[4]) & 1) != 0;
    for (int t = tid; t < (WSMG_SKIP(512) ? 0 : fan_cells); t += NT) {
      F4 v;
#pragma unroll
      for (int ch = 0; ch < SLAB; ++ch) {
        int32_t k = Pk[ch * npp + t];
        if (t == 0 && inv && k < sentinel_key) k = sentinel_key;   // invalid pixels write -1e16 to cell 0 (:207-212)
        v.v[ch] = ch < nch ? finish_cell(k) : 0.0f;
      }
      X[1 + t] = v;
    }
    WSMG_SYNC();
    for (int t = tid; t < (WSMG_SKIP(512) ? 0 : fan_cells); t += NT) Pf[1 + t] = X[1 + t];
  } else {
    // stage API: load the projection (zero outside the fan by construction)
    const float* src = p.proj_in + ((size_t)b * C + c0) * EE;
    for (int y = 0; y < g.fan_rows; ++y) {
      int xs = fan_x_lo(y), w = fan_row_width(y, E);
      int base = fanrow[y].a + xs;
      for (int t = tid; t < w; t += NT) {
        F4 v;
#pragma unroll
        for (int ch = 0; ch < SLAB; ++ch) v.v[ch] = ch < nch ? src[(size_t)ch * EE + y * E + xs + t] : 0.0f;
        Pf[base + t] = v;
      }
    }
  }
  WSMG_SYNC();

  // fan tap: row info -> index (0 = zero cell when the row or the column is outside the fan)
  auto fan_idx = [&](const I2 fr, int x) -> int {
    int xs = fr.b & 0xFFFF, xe = fr.b >> 16;
    return (x >= xs && x <= xe) ? fr.a + x : 0;
  };

  if (p.proj_out != nullptr) {
    float* dst = p.proj_out + ((size_t)b * C + c0) * EE;
    for (int t = tid; t < EE; t += NT) {
      int y = fd_div(t, E, mE), x = t - y * E;
      F4 v = Pf[fan_idx(fanrow[y], x)];
#pragma unroll
      for (int ch = 0; ch < SLAB; ++ch)
        if (ch < nch) dst[(size_t)ch * EE + t] = v.v[ch];
    }
  }
  if (p.stop_after_scatter) return;

  // ---- phase 2: R = rotate(P, -compass) into X (rgb_mapping.py:37 -> :267 -> :239-250) ---------
  float cs = scal[0], sn = scal[1];
  // Cells none of whose taps falls inside the fan (~70 % of the grid) are exact zeros: store and move on.
  // The per-row extent of the other cells lets the fuse step skip window cells that only see zeros.
  const int nslots = tile_slots(E);
  for (int s0 = 0; s0 < (WSMG_SKIP(2) ? 0 : nslots); s0 += NT) {
    const int slot = s0 + tid;
    bool hit = false;
    int i = 0, j = 0;
    bool in_tile = slot < nslots && tile_cell(slot, E, &i, &j, mT);
    if (in_tile) {
      const int hr = hitrow[i];
      if (j < (hr & 0xFFFF) || j >= (hr >> 16)) { X[1 + i * E + j] = f4_zero(); in_tile = false; }   // cannot see the fan
    }
    if (in_tile) {
      const int t = i * E + j;
      float ix, iy;
      rot_coords(baseE[j], baseE[i], cs, sn, half_e, &ix, &iy);
      Tap1D tx = make_tap(ix), ty = make_tap(iy);
      int y0 = ty.i0, y1 = ty.i0 + 1;
      int ia = 0, ib = 0, ic = 0, id = 0;
      if (y1 >= 0 && y0 < g.fan_rows) {                      // half of the grid samples rows the fan does not have
        I2 r0 = fanrow[(unsigned)y0 < (unsigned)E ? y0 : E];
        I2 r1 = fanrow[(unsigned)y1 < (unsigned)E ? y1 : E];
        ia = fan_idx(r0, tx.i0); ib = fan_idx(r0, tx.i0 + 1); ic = fan_idx(r1, tx.i0); id = fan_idx(r1, tx.i0 + 1);
      }
      hit = (ia | ib | ic | id) != 0;
      if (hit) {
        Weights w = make_weights(tx.w1, ty.w1);
        X[1 + t] = blend_f4(tap(Pf, ia), tap(Pf, ib), tap(Pf, ic), tap(Pf, id), w);
      } else {
        X[1 + t] = f4_zero();
      }
    }
#if defined(__CUDACC__)
    // per-row extents of the warp's 8 x 4 tile: lane r (0..3) owns row r.  Row r's cells sit in lanes
    // {a_r..a_r+3} (columns 0-3) and {a_r+8..a_r+11} (columns 4-7), a_r = 0, 4, 16, 20.
    const unsigned m = __ballot_sync(0xFFFFFFFFu, hit);
    if (m != 0u) {
      const int lane = tid & 31;
      if (lane < 4) {
        const int a = (lane & 1) * 4 + (lane >> 1) * 16;
        const unsigned mr = ((m >> a) & 0xFu) | (((m >> (a + 8)) & 0xFu) << 4);
        if (mr != 0u) {
          const int tile = (s0 + tid) >> 5, tiles_x = (E + 7) >> 3;
          const int band = fd_div(tile, tiles_x, mT), j0 = 8 * (tile - band * tiles_x), row = 4 * band + lane;
          atomicMin(&ext[row].a, j0 + __ffs(mr) - 1);
          atomicMax(&ext[row].b, j0 + 31 - __clz(mr));
        }
      }
    }
#else
    if (hit) { if (j < ext[i].a) ext[i].a = j; if (j > ext[i].b) ext[i].b = j; }
#endif
  }
  WSMG_SYNC();                     // extents complete

  // ---- phase 3 tables: the two (separable) translations (rgb_mapping.py:45-53, 57-65) ----------
  // (NEG also marks a second tap whose weight is exactly 0 -- two thirds of the rows / columns)
  // colT[vv] = {x0 | NEG, x0+1 | NEG, bits(wx), NEG if the column is outside the map else (leftmost tap col) | (rightmost tap col + 1) << 10}
  // rowT[uu] = {1 + y0*E | NEG, 1 + (y0+1)*E | NEG, bits(wy), 1 + slot(uu)*WWP if the row is inside the map else NEG}
  // bXT[q]   = {col0 | NEG, col1 | NEG, bits(wx), 0}
  // bYT[p]   = {1 + slot(row0)*WWP | NEG, 1 + slot(row1)*WWP | NEG, bits(wy), 0}     (slot(r) = (r + S0) % RR)
  for (int t = tid; t < (WSMG_SKIP(1024) ? 0 : WW); t += NT) {
    int v = v0 + t, u = u0 + t;
    I4 ct; ct.a = ct.b = NEG; ct.c = 0; ct.d = NEG;
    if ((unsigned)v < (unsigned)G) {    // canvas column sampled by global column v, relative to the pasted ego grid
      Tap1D tp = make_tap(unnormalize(base_coord(v, G) + (-qx), half_g));
      int x0 = tp.i0 - paste_lo;
      ct.a = (unsigned)x0 < (unsigned)E ? x0 : NEG;
      ct.b = ((unsigned)(x0 + 1) < (unsigned)E && tp.w1 != 0.0f) ? x0 + 1 : NEG;     // weight exactly 0: skip the tap
      ct.c = as_int(tp.w1);
      const int cl = ct.a >= 0 ? ct.a : (ct.b >= 0 ? ct.b : E + 1);     // leftmost / rightmost source column actually read
      const int ch = ct.b >= 0 ? ct.b : ct.a;                           // (ct.a < 0 and ct.b < 0: none, ch + 1 <= 0)
      ct.d = cl | ((ch >= 0 ? ch + 1 : 0) << 10);
    }
    colT[t] = ct;
    I4 rt; rt.a = rt.b = NEG; rt.c = 0; rt.d = NEG;
    if ((unsigned)u < (unsigned)G) {
      Tap1D tp = make_tap(unnormalize(base_coord(u, G) + (-qy), half_g));
      int y0 = tp.i0 - paste_lo;
      rt.a = (unsigned)y0 < (unsigned)E ? 1 + y0 * E : NEG;
      rt.b = ((unsigned)(y0 + 1) < (unsigned)E && tp.w1 != 0.0f) ? 1 + (y0 + 1) * E : NEG;
      rt.c = as_int(tp.w1); rt.d = 1 + ring_slot(t) * WWP;
    }
    rowT[t] = rt;
    {  // columns of R that can contribute to window row t: union of the extents of its (up to) two source rows
      I2 e; e.a = E; e.b = -1;
      if (rt.a > 0) { I2 s0e = ext[fd_div(rt.a - 1, E, mE)]; e = s0e; }
      if (rt.b > 0) { I2 s1e = ext[fd_div(rt.b - 1, E, mE)]; e.a = s1e.a < e.a ? s1e.a : e.a; e.b = s1e.b > e.b ? s1e.b : e.b; }
      rowE[t] = e;
    }
  }
  for (int t = tid; t < (WSMG_SKIP(1024) ? 0 : E); t += NT) {   // global column / row sampled by crop cell t, relative to the window
    Tap1D tp = make_tap(unnormalize(base_coord(t + paste_lo, G) + qx, half_g));
    int cx = tp.i0 - v0;
    I4 bx; bx.a = (unsigned)cx < (unsigned)WW ? cx : NEG;
    bx.b = ((unsigned)(cx + 1) < (unsigned)WW && tp.w1 != 0.0f) ? cx + 1 : NEG;
    bx.c = as_int(tp.w1); bx.d = 0;
    bXT[t] = bx;
    Tap1D tq = make_tap(unnormalize(base_coord(t + paste_lo, G) + qy, half_g));
    int ry = tq.i0 - u0;
    I4 by; by.a = (unsigned)ry < (unsigned)WW ? 1 + ring_slot(ry) * WWP : NEG;
    by.b = ((unsigned)(ry + 1) < (unsigned)WW && tq.w1 != 0.0f) ? 1 + ring_slot(ry + 1) * WWP : NEG;
    by.c = as_int(tq.w1); by.d = 0;
    bYT[t] = by;
  }
  if (TMA) fence_proxy_async();    // generic writes to the key planes precede TMA writes to the same bytes
  WSMG_SYNC();                     // R complete, fan dead: the key planes may now be overwritten by ring rows
  if (NB > 1) prefetch_band(1);

  // ---- phase 3: banded translate + max-fuse (:53-56) and translate back + crop (:64-69) -------
  // Trip k fuses band k (F rows [kB, kB+B)) in place in the ring and, independently, back-translates
  // the crop rows whose F rows were finished by trip k-1: B row p reads F rows p..p+2 and overwrites
  // X row p, which only F rows p..p+2 read -- so rows p <= done-3 are safe.  One barrier per trip.
  int p_lo = 0;
  auto fuse_band = [&](int k) {                              // 3a: F rows of band k, in place in the ring
    for (int t = tid; t < BAND * WW; t += LT) {
      int rr = fd_div(t, WW, mWW), vv = t - rr * WW, uu = k * BAND + rr;
      if (uu >= WW) continue;
      const I4 ct = colT[vv], rt = rowT[uu];
      const int cell = rt.d + vv;
      const I2 e = rowE[uu];                                  // R columns that are not identically zero for this row
      // row and column inside the map, and some tap column inside the extent; otherwise T == 0 and
      // F = max(G, 0) = G: nothing to do
      if (rt.d > 0 && ct.d >= 0 && (ct.d >> 10) > e.a && (ct.d & 1023) <= e.b) {
        Weights w = make_weights(as_float(ct.c), as_float(rt.c));
        F4 a = tap(X, rt.a + ct.a), bb = tap(X, rt.a + ct.b), c = tap(X, rt.b + ct.a), d = tap(X, rt.b + ct.b);
        F4 tv = blend_f4(a, bb, c, d, w);
        F4* cellp = ring + cell;
        const F4 old = *cellp;
        F4 f;
        bool changed = false;
#pragma unroll
        for (int ch = 0; ch < SLAB; ++ch) {
          f.v[ch] = fmaxf(old.v[ch], tv.v[ch]);
          changed |= as_int(f.v[ch]) != as_int(old.v[ch]);
        }
        if (changed) {             // cells the observation does not raise keep their bytes: no store, no DRAM write
          *cellp = f;
          if (TMA) fence_proxy_async();   // writer-side: this generic write precedes the TMA load that later recycles the row
          float* dst = gmap_b + ((size_t)(u0 + uu) * G + (v0 + vv)) * C;
          if (VEC) {
            *reinterpret_cast<F4*>(dst) = f;
          } else {
#pragma unroll
            for (int ch = 0; ch < SLAB; ++ch)
              if (ch < nch) dst[ch] = f.v[ch];
          }
        }
      }
    }
  };
  auto crop_rows = [&](int k) {                              // 3b: B rows whose F rows were finished before trip k
    int done = k * BAND < WW ? k * BAND : WW;
    int p_hi = done - 2 > p_lo ? done - 2 : p_lo;
    for (int t = tid; t < (p_hi - p_lo) * E; t += LT) {
      int dr = fd_div(t, E, mE), q = t - dr * E, pr = p_lo + dr;
      const I4 bx = bXT[q], by = bYT[pr];
      Weights w = make_weights(as_float(bx.c), as_float(by.c));
      F4 a = tap(ring, by.a + bx.a), bb = tap(ring, by.a + bx.b), c = tap(ring, by.b + bx.a), d = tap(ring, by.b + bx.b);
      X[1 + pr * E + q] = blend_f4(a, bb, c, d, w);
    }
    p_lo = p_hi;
  };
  if (TMA && tid >= LT) {
    // no cell of the loop is ours: meet the others at the barrier behind it
  } else
  for (int k = 0; k <= (WSMG_SKIP(4) ? -1 : NB); ++k) {
    if (!TMA && k < NB) {
      if (k + 1 < NB) async_wait<1>(); else async_wait<0>();
    }
    if (TMA) group_sync(1, LT); else WSMG_SYNC();           // (the cp.async build prefetches with every thread)
    if (TMA) {
      // ring rows of band k+2 were last read by trip k-1's crop (RR >= 4*BAND + 2) and last written by an
      // earlier fuse (fenced below): the barrier above makes them free for the TMA unit to overwrite
      if ((unsigned)tma_lane < 32u && k + 2 < NB) prefetch_band(k + 2);
      if (k >= 1 && !WSMG_SKIP(16)) crop_rows(k);    // needs only finished F rows: runs while band k is still landing
      if (k < NB) {
        if (!WSMG_SKIP(64)) mbar_wait(&bars[k], 0);
        if (!WSMG_SKIP(32)) fuse_band(k);
      }
    } else {
      if (k + 2 < NB) prefetch_band(k + 2);
      else async_commit();         // keep one group per trip so wait<1> stays exact
      if (k < NB) fuse_band(k);
      if (k >= 1) crop_rows(k);
    }
  }
  WSMG_SYNC();

  // ---- phase 4: ego = rotate(B, +compass), NCHW out (rgb_mapping.py:70) ----------------------
  cs = scal[2]; sn = scal[3];
  float* ego_b = p.ego + ((size_t)b * C + c0) * EE;
  for (int slot = tid; slot < (WSMG_SKIP(8) ? 0 : nslots); slot += NT) {
    int i, j;
    if (!tile_cell(slot, E, &i, &j, mT)) continue;
    const int t = i * E + j;
    float ix, iy;
    rot_coords(baseE[j], baseE[i], cs, sn, half_e, &ix, &iy);
    Tap1D tx = make_tap(ix), ty = make_tap(iy);
    Weights w = make_weights(tx.w1, ty.w1);
    int x0 = tx.i0, y0 = ty.i0;
    int cx0 = (unsigned)x0 < (unsigned)E ? x0 : NEG, cx1 = (unsigned)(x0 + 1) < (unsigned)E ? x0 + 1 : NEG;
    int ry0 = (unsigned)y0 < (unsigned)E ? 1 + y0 * E : NEG, ry1 = (unsigned)(y0 + 1) < (unsigned)E ? 1 + (y0 + 1) * E : NEG;
    F4 a = X[imax0(ry0 + cx0)], bb = X[imax0(ry0 + cx1)], c = X[imax0(ry1 + cx0)], d = X[imax0(ry1 + cx1)];
    F4 r = blend_f4(a, bb, c, d, w);
#pragma unroll
    for (int ch = 0; ch < SLAB; ++ch)
      if (ch < nch) st_stream(ego_b + (size_t)ch * EE + t, r.v[ch]);
    if (p.ego_half != nullptr) {
      uint16_t* half_b = p.ego_half + ((size_t)b * C + c0) * EE;
#pragma unroll
      for (int ch = 0; ch < SLAB; ++ch)
        if (ch < nch) half_b[(size_t)ch * EE + t] = f2half_bits(r.v[ch]);
    }
  }
}

}  // namespace wsmg
