// libwsmg.so -- WS-MGMap per-step map update for B200 (sm_100a).  C ABI in include/wsmg.h.
//
// Two launches per step, both on the caller's stream, the second a programmatic dependent launch of the first:
//   k_cells   fused unproject + height-band test + bin + index  (rgb_mapping.py:153-176, 188-217); its trailing blocks
//             apply the episode-reset mask to the NHWC map (rgb_mapping.py:35) and evaluate the per-env rotation
//             sines / cosines and column bounds (k_reset does the same for the stage entry points)
//   k_fused   one CTA per (env, 4-channel slab): shared-memory scatter-max, rotate, translate,
//             max-fuse into the map window, translate back, crop, rotate -> NCHW ego map
//             (rgb_mapping.py:210-232, 37-70).  Body in wsmg_body.h.
// No tensor cores: the path is gather / scatter-max / bilinear streaming (~0.3 flop/byte).
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include <vector>

#include "wsmg_body.h"
#include "wsmg_host.h"

namespace wsmg {

constexpr int CELLS_THREADS = 256;

// ------------------------------------------------------------------ per-env preparation shared by k_reset and k_cells
// In the whole-step launch (`pre.cell_blocks > 0`) the grid carries, per env, 1 + PRE_RESET_BLOCKS more blocks that do what
// k_reset does for the stage entry points -- the rotations' cos / sin and column bounds, the bad-slot check, the
// episode-reset mask -- so that a step is two launches, and every block leaves its flags in a word of its own
// (block_flags[b][blk], plain store: nothing to clear beforehand); k_fused ORs the env's words.
constexpr int PRE_RESET_BLOCKS = 16;
struct PreArgs {
  int cell_blocks = 0;             // > 0: whole-step launch with the extra roles below
  uint32_t* block_flags = nullptr; // [bs][cell_blocks + 1]
  float* gmap = nullptr; const float* mask = nullptr; size_t per_env = 0; const int32_t* env_slots = nullptr;
  int32_t* row_bounds = nullptr; float* env_trig = nullptr; const float* compass = nullptr; const float* trig = nullptr;
  int n_maps = 0;
};

__device__ __forceinline__ void reset_map_rows(float* __restrict__ base, float m, size_t per_env, int chunk, int chunks) {
  const size_t stride = (size_t)chunks * blockDim.x;
  size_t i = (size_t)chunk * blockDim.x + threadIdx.x;
  if ((per_env & 3) == 0) {
    float4* b4 = reinterpret_cast<float4*>(base);
    const size_t n4 = per_env >> 2;
    if (m == 0.0f) {
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      for (; i < n4; i += stride) b4[i] = z;
    } else {
      for (; i < n4; i += stride) {
        float4 v = b4[i];
        v.x *= m; v.y *= m; v.z *= m; v.w *= m;
        b4[i] = v;
      }
    }
  } else {
    for (; i < per_env; i += stride) base[i] = (m == 0.0f) ? 0.0f : base[i] * m;
  }
}

// both rotations' cos / sin exactly as k_fused would evaluate them, and the first rotation's per-row column bounds
__device__ __forceinline__ void env_rotation_setup(int b, const Geo& g, const float* __restrict__ compass, const float* __restrict__ trig,
                                                   int32_t* __restrict__ row_bounds, float* __restrict__ env_trig) {
  float cs, sn;
  if (trig != nullptr) { cs = trig[4 * b + 0]; sn = trig[4 * b + 1]; }
  else { const float h = -compass[b]; sn = sinf(h); cs = cosf(h); }
  for (int t = threadIdx.x; t < g.E; t += blockDim.x) row_bounds[(size_t)b * g.E + t] = rot_row_bounds(g, cs, sn, t);
  if (env_trig != nullptr && threadIdx.x == 0) {
    float cs2, sn2;
    if (trig != nullptr) { cs2 = trig[4 * b + 2]; sn2 = trig[4 * b + 3]; }
    else { const float h = compass[b]; sn2 = sinf(h); cs2 = cosf(h); }
    env_trig[4 * b + 0] = cs; env_trig[4 * b + 1] = sn; env_trig[4 * b + 2] = cs2; env_trig[4 * b + 3] = sn2;
  }
}

// ------------------------------------------------------------------ k_reset
// Per-env preparation for the STAGE entry points (the whole step folds it into the k_cells launch), grid (bs, chunks):
//   * full_global_map[:bs] *= masks (rgb_mapping.py:35); mask == 1 (the steady state) touches nothing;
//   * clears the env flags (k_cells, the next launch, sets them);
//   * once per env: both rotations' cos / sin exactly as the rotations will use them, and the per-row column bounds
//     outside which the first rotation cannot see the fan (rot_row_bounds, wsmg_body.h);
//   * flags an env slot that is not a row of the map tensor (the frame is then skipped everywhere).
__global__ void __launch_bounds__(256) k_reset(float* __restrict__ gmap, const float* __restrict__ mask,
                                               size_t per_env, uint32_t* __restrict__ env_flags,
                                               const int32_t* __restrict__ env_slots, int32_t* __restrict__ row_bounds,
                                               float* __restrict__ env_trig, const float* __restrict__ compass,
                                               const float* __restrict__ trig, uint32_t* __restrict__ status, int n_maps, Geo g) {
  const int b = blockIdx.x, chunk = blockIdx.y;
  const int mrow = env_slots != nullptr ? env_slots[b] : b;
  const bool bad_slot = gmap != nullptr && (unsigned)mrow >= (unsigned)n_maps;
  if (chunk == 0 && threadIdx.x == 0) {
    if (env_flags != nullptr) env_flags[b] = bad_slot ? WSMG_FLAG_BAD_SLOT : 0u;
    if (bad_slot && status != nullptr) status[1] = 1u;
  }
  if (row_bounds != nullptr && chunk == gridDim.y - 1) env_rotation_setup(b, g, compass, trig, row_bounds, env_trig);
  if (gmap == nullptr || bad_slot) return;
  const float m = mask[b];
  if (m != 1.0f) reset_map_rows(gmap + (size_t)mrow * per_env, m, per_env, chunk, gridDim.y);
}

// ------------------------------------------------------------------ k_cells
// Fused unproject + height-band test + bin + index.  Each thread owns four consecutive sampled pixels
// of the Hf x Wf frame (one 8-byte store of packed fan codes).  The depth rows a block needs are one
// contiguous span of the depth image: a single TMA bulk copy (cp.async.bulk + mbarrier complete_tx)
// stages them in shared memory while the per-column pinhole table is being built; the nearest-
// neighbour subsampling (rgb_mapping.py:188-196) then gathers from shared memory.  Also emits the
// reference-shaped (linear index, invalid) pair for the stage API and the per-env "some pixel does not
// write" flag (those pixels send the sentinel to cell 0, rgb_mapping.py:207-212).
constexpr int CELLS_PX = 4;             // pixels per packed code word
constexpr int CELLS_GROUPS = 2;         // code words per thread: the per-block tables are built once per 2048 pixels
constexpr int CELLS_MAX_W = 1024;
// STAGE_API: also write the reference-shaped (lin, invalid) arrays (stage entry point only; the step never does).
template <bool STAGE_API>
__global__ void __launch_bounds__(CELLS_THREADS) k_cells(const float* __restrict__ depth, uint16_t* __restrict__ codes,
                                                          int32_t* __restrict__ lin, uint8_t* __restrict__ invalid,
                                                          uint32_t* __restrict__ env_flags, uint32_t* __restrict__ status, Geo g, int stage_rows,
                                                          const PreArgs pre) {
  // k_fused is launched with programmatic stream serialization behind this kernel: let its CTAs be scheduled as soon as
  // every block of this grid has started (they wait for this grid's completion before touching what it writes)
  asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
  if (pre.cell_blocks > 0 && (int)blockIdx.y >= pre.cell_blocks) {
    const int b = blockIdx.x, role = (int)blockIdx.y - pre.cell_blocks;
    const int mrow = pre.env_slots != nullptr ? pre.env_slots[b] : b;
    const bool bad_slot = (unsigned)mrow >= (unsigned)pre.n_maps;
    if (role == 0) {
      if (threadIdx.x == 0) {
        pre.block_flags[(size_t)b * (pre.cell_blocks + 1) + pre.cell_blocks] = bad_slot ? WSMG_FLAG_BAD_SLOT : 0u;
        if (bad_slot && status != nullptr) status[1] = 1u;
      }
      env_rotation_setup(b, g, pre.compass, pre.trig, pre.row_bounds, pre.env_trig);
    } else if (!bad_slot) {
      const float m = pre.mask[b];
      if (m != 1.0f) reset_map_rows(pre.gmap + (size_t)mrow * pre.per_env, m, pre.per_env, role - 1, PRE_RESET_BLOCKS);
    }
    return;
  }
  extern __shared__ __align__(128) unsigned char cells_smem[];
  __shared__ int rowoff[160];
  __shared__ __align__(16) int col_src[CELLS_MAX_W];
  __shared__ __align__(16) float col_xx[CELLS_MAX_W];
  __shared__ __align__(8) uint64_t bar;
  float* drows = reinterpret_cast<float*>(cells_smem);       // [stage_rows][Wd] when stage_rows > 0
  const int b = blockIdx.x;

  const int HW = g.Hf * g.Wf;
  const int per_block = CELLS_THREADS * CELLS_PX * CELLS_GROUPS;
  const int t_first = blockIdx.y * per_block;
  const int t_last = (t_first + per_block < HW ? t_first + per_block : HW) - 1;
  const float* depth_b = depth + (size_t)b * g.Hd * g.Wd;
  const int r_first = sample_index(g, t_first / g.Wf);
  const int n_rows = sample_index(g, t_last / g.Wf) - r_first + 1;
  const bool staged = stage_rows > 0 && n_rows <= stage_rows;
  if (staged && threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_init_fence();
    const unsigned bytes = (unsigned)(n_rows * g.Wd * 4);
    mbar_expect_tx(&bar, bytes);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(saddr(drows)), "l"(depth_b + (size_t)r_first * g.Wd), "r"(bytes), "r"(saddr(&bar)) : "memory");
  }
  for (int t = threadIdx.x; t < g.fan_rows; t += blockDim.x) rowoff[t] = fan_row_offset(t, g.E);
  for (int j = threadIdx.x; j < g.Wf; j += blockDim.x) {
    int c = sample_index(g, j);
    col_src[j] = c;
    col_xx[j] = pinhole_xx(g, c);
  }
  __syncthreads();
  if (staged) mbar_wait(&bar, 0);
  bool any_bad = false, any_outlier = false;
#pragma unroll
  for (int grp = 0; grp < CELLS_GROUPS; ++grp) {
    const int t0 = t_first + (grp * CELLS_THREADS + threadIdx.x) * CELLS_PX;   // a warp's words stay consecutive
    int i = t0 / g.Wf, j = t0 - i * g.Wf;
    int r = sample_index(g, i);
    float yy = pinhole_yy(g, r);
    uint32_t code[CELLS_PX];
    auto one_pixel = [&](int px, int t, float dval, float xx) {
      int x, y;
      const bool ok = unproject_depth(g, dval, xx, yy, &x, &y);
      any_bad |= !ok;
      code[px] = CODE_INVALID;
      if (ok) {
        if (y < g.fan_rows && x >= fan_x_lo(y) && x <= fan_x_hi(y, g.E)) code[px] = (uint32_t)(rowoff[y] + x - fan_x_lo(y));
        else { code[px] = CODE_OUTLIER; any_outlier = true; }
      }
      if (STAGE_API) {
        if (lin != nullptr) lin[(size_t)b * HW + t] = y * g.E + x;
        if (invalid != nullptr) invalid[(size_t)b * HW + t] = ok ? 0 : 1;
      }
    };
    if ((g.Wf & 3) == 0 && t0 + CELLS_PX <= HW) {
      // the four pixels share a row: one row pointer, the column tables as two 16-byte reads
      const int4 cs = *reinterpret_cast<const int4*>(&col_src[j]);
      const float4 cx = *reinterpret_cast<const float4*>(&col_xx[j]);
      float d0, d1, d2, d3;
      if (staged) {                                            // (kept apart so that the staged reads are LDS, not generic loads)
        const float* row = drows + (r - r_first) * g.Wd;
        d0 = row[cs.x]; d1 = row[cs.y]; d2 = row[cs.z]; d3 = row[cs.w];
      } else {
        const float* row = depth_b + (size_t)r * g.Wd;
        d0 = __ldg(row + cs.x); d1 = __ldg(row + cs.y); d2 = __ldg(row + cs.z); d3 = __ldg(row + cs.w);
      }
      one_pixel(0, t0, d0, cx.x); one_pixel(1, t0 + 1, d1, cx.y); one_pixel(2, t0 + 2, d2, cx.z); one_pixel(3, t0 + 3, d3, cx.w);
    } else {
  #pragma unroll
      for (int px = 0; px < CELLS_PX; ++px) {
        const int t = t0 + px;
        code[px] = CODE_INVALID;
        if (t < HW) {
          const float dval = staged ? drows[(r - r_first) * g.Wd + col_src[j]] : depth_b[(size_t)r * g.Wd + col_src[j]];
          one_pixel(px, t, dval, col_xx[j]);
          if (++j == g.Wf) { j = 0; ++i; r = sample_index(g, i); yy = pinhole_yy(g, r); }
        }
      }
    }
    if (codes != nullptr && t0 < HW) {       // HW % 4 == 0 (validated): the four codes are one aligned 8-byte word
      uint2 w; w.x = code[0] | (code[1] << 16); w.y = code[2] | (code[3] << 16);
      *reinterpret_cast<uint2*>(codes + (size_t)b * HW + t0) = w;
    }
  }
  const unsigned bad = __ballot_sync(0xFFFFFFFFu, any_bad);
  const unsigned outl = __ballot_sync(0xFFFFFFFFu, any_outlier);
  const uint32_t word = (bad ? WSMG_FLAG_INVALID_PIXEL : 0u) | (outl ? WSMG_FLAG_OUTSIDE_FAN : 0u);
  if (pre.cell_blocks > 0) {                                 // whole-step launch: this block's own word
    __shared__ uint32_t blk_word;
    if (threadIdx.x == 0) blk_word = 0u;
    __syncthreads();
    if (word != 0u && (threadIdx.x & 31) == 0) atomicOr(&blk_word, word);
    __syncthreads();
    if (threadIdx.x == 0) pre.block_flags[(size_t)b * (pre.cell_blocks + 1) + blockIdx.y] = blk_word;
  } else if (env_flags != nullptr && word != 0u && (threadIdx.x & 31) == 0) {
    atomicOr(env_flags + b, word);
  }
  if (status != nullptr && outl != 0u && (threadIdx.x & 31) == 0) status[0] = 1u;   // sticky, polled by the host at its next call
}

// ------------------------------------------------------------------ k_fused
template <int CE, int CG, int CHW, bool VEC, bool TMA, int FEAT>
__global__ void __launch_bounds__(FUSED_NT, 1) k_fused(const __grid_constant__ FusedParams p) {
  // One CTA per (env, slab).  (A persistent loop over the items was measured twice -- one 1024-thread CTA per SM in
  // round 1, two 512-thread CTAs per SM in round 2 -- and is not faster: the hardware already overlaps CTA launch with
  // the previous CTA's tail.)
  extern __shared__ __align__(1024) unsigned char smem[];
  fused_body<FUSED_NT, CE, CG, CHW, VEC, TMA, FEAT>(p, blockIdx.x, smem, threadIdx.x);
}

// ------------------------------------------------------------------ k_semcrop
// The ground-truth semantic map sensor for a batch of envs: one thread per output cell, two chained nearest
// resamplings collapsed into one gather (sensors.py:403-410).  4*half^2 cells per env: latency-bound, tiny.
__global__ void __launch_bounds__(256) k_semcrop(const float* __restrict__ maps, const float* __restrict__ pose,
                                                 const float* __restrict__ trig, const int32_t* __restrict__ map_index,
                                                 long long* __restrict__ out, int n_maps, int S, int half, int origin) {
  const int b = blockIdx.y, side = 2 * half;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= side * side) return;
  float cs, sn;
  if (trig != nullptr) { cs = trig[2 * b]; sn = trig[2 * b + 1]; }
  else { const float h = pose[3 * b + 2]; cs = cosf(h); sn = sinf(h); }
  const int i = t / side, j = t - i * side;
  const int idx = semmap_source_index(origin - side + i, origin - side + j, S, cs, sn, pose[3 * b], pose[3 * b + 1]);
  const int m = map_index != nullptr ? map_index[b] : b;
  long long v = 0;
  if (idx >= 0 && (unsigned)m < (unsigned)n_maps) v = (long long)maps[(size_t)m * S * S + idx];
  out[(size_t)b * side * side + t] = v;
}

// ------------------------------------------------------------------ host glue
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Geometry constants and the shared-memory plan of the last dims seen by this host thread (a trainer calls with the
// same shapes step after step; make_geo evaluates a tangent and sums the fan, make_plan sizes the ring).
struct GeoPlan { Geo g; SmemPlan sp; };
static const GeoPlan& geo_plan(const wsmg_dims* d) {
  static thread_local wsmg_dims last{};
  static thread_local GeoPlan gp;
  static thread_local bool valid = false;
  const bool same = valid && d->C == last.C && d->Hf == last.Hf && d->Wf == last.Wf && d->Hd == last.Hd && d->Wd == last.Wd &&
                    d->E == last.E && d->G == last.G && d->resolution == last.resolution && d->C_in == last.C_in && d->feat_nhwc == last.feat_nhwc;   // (bs, n_maps: not geometry)
  if (!same) {
    gp.g = make_geo(d);
    gp.sp = make_plan(gp.g);
    last = *d;
    valid = true;
  }
  return gp;
}

struct ResetArgs {      // optional outputs of k_reset beyond the mask reset
  uint32_t* env_flags = nullptr; const int32_t* env_slots = nullptr; int32_t* row_bounds = nullptr;
  float* env_trig = nullptr; const float* compass = nullptr; const float* trig = nullptr; uint32_t* status = nullptr;
};

static int launch_reset(float* gmap, const float* mask, const wsmg_dims* d, cudaStream_t s, const ResetArgs& a) {
  const size_t per_env = (size_t)d->G * d->G * d->C;
  dim3 grid(d->bs, gmap ? 16 : 1);
  k_reset<<<grid, 256, 0, s>>>(gmap, mask, per_env, a.env_flags, a.env_slots, a.row_bounds, a.env_trig, a.compass, a.trig,
                              a.status, d->n_maps, geo_plan(d).g);
  return (int)cudaGetLastError();
}

static int cell_blocks_of(const Geo& g) {
  const int per_block = CELLS_THREADS * CELLS_PX * CELLS_GROUPS;
  return (g.Hf * g.Wf + per_block - 1) / per_block;
}

// `pre` != nullptr: the whole-step launch (extra per-env blocks, per-block flag words; pre->cell_blocks is filled here).
static int launch_cells(const float* depth, uint16_t* codes, int32_t* lin, uint8_t* invalid, uint32_t* env_flags,
                        uint32_t* status, const Geo& g, int bs, cudaStream_t s, PreArgs* pre = nullptr) {
  if (g.fan_rows > 160 || g.Wf > CELLS_MAX_W) return WSMG_E_DIMS;
  const int per_block = CELLS_THREADS * CELLS_PX * CELLS_GROUPS;
  const int cell_blocks = cell_blocks_of(g);
  PreArgs pa;
  if (pre != nullptr) { pre->cell_blocks = cell_blocks; pa = *pre; }
  dim3 grid(bs, cell_blocks + (pre != nullptr ? 1 + PRE_RESET_BLOCKS : 0));
  if (grid.y > 65535u) return WSMG_E_DIMS;
  // depth rows one block can touch: its sampled rows (per_block / Wf + 2) times the subsampling ratio, + 1
  int stage_rows = (int)((per_block / g.Wf + 2) * (double)g.Hd / g.Hf) + 2;
  size_t smem = (size_t)stage_rows * g.Wd * 4;
  const bool bulk_ok = (g.Wd % 4) == 0 && (((size_t)g.Hd * g.Wd) % 4) == 0 && (reinterpret_cast<uintptr_t>(depth) & 15u) == 0;
  if (smem > 32 * 1024 || !bulk_ok) { stage_rows = 0; smem = 0; }     // fall back to direct global gathers
  if (lin != nullptr || invalid != nullptr) k_cells<true><<<grid, CELLS_THREADS, smem, s>>>(depth, codes, lin, invalid, env_flags, status, g, stage_rows, pa);
  else k_cells<false><<<grid, CELLS_THREADS, smem, s>>>(depth, codes, lin, invalid, env_flags, status, g, stage_rows, pa);
  return (int)cudaGetLastError();
}

// CUtensorMap of the caller's NHWC map [n_maps, G, G, C] with a box of {4 ch, WWP cols, TMA_ROWS rows, 1}.
// cuTensorMapEncodeTiled is resolved through the runtime (no link-time dependency on libcuda).  The last encoding is
// kept per host thread: a trainer calls with the same map tensor step after step.
static int encode_window_map(TensorMapBlob* out, float* gmap, int n_maps, const Geo& g, int wwp) {
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn fn = nullptr;      // process-wide constant once resolved
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e != cudaSuccess) return (int)e;
    if (ptr == nullptr || qres != cudaDriverEntryPointSuccess) return (int)cudaErrorNotSupported;
    fn = (encode_fn)ptr;
  }
  static_assert(sizeof(CUtensorMap) <= sizeof(TensorMapBlob), "tensor map blob too small");
  struct Key { float* gmap; int n_maps, G, C, wwp; };
  static thread_local Key last_key = {nullptr, 0, 0, 0, 0};
  static thread_local TensorMapBlob last_blob;
  if (last_key.gmap == gmap && last_key.n_maps == n_maps && last_key.G == g.G && last_key.C == g.C && last_key.wwp == wwp) {
    *out = last_blob;
    return 0;
  }
  CUtensorMap tm;
  const cuuint64_t dims[4] = {(cuuint64_t)g.C, (cuuint64_t)g.G, (cuuint64_t)g.G, (cuuint64_t)n_maps};
  const cuuint64_t strides[3] = {(cuuint64_t)g.C * 4, (cuuint64_t)g.G * g.C * 4, (cuuint64_t)g.G * g.G * g.C * 4};
  const cuuint32_t box[4] = {4, (cuuint32_t)wwp, (cuuint32_t)TMA_ROWS, 1};   // see prefetch_band
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, gmap, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return (int)cudaErrorInvalidValue;
  memcpy(out->bytes, &tm, sizeof(tm));
  last_key = Key{gmap, n_maps, g.G, g.C, wwp};
  last_blob = *out;
  return 0;
}

// Per-device constants and per-process switches, looked up once.
struct DeviceInfo { int sms = 0, max_optin = 0, dev = 0; };
static int device_info(DeviceInfo* out) {
  static DeviceInfo cache[64];
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  if (dev < 0 || dev >= 64) return (int)cudaErrorInvalidDevice;
  if (cache[dev].sms == 0) {
    DeviceInfo di;
    if ((e = cudaDeviceGetAttribute(&di.max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)) != cudaSuccess) return (int)e;
    if ((e = cudaDeviceGetAttribute(&di.sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return (int)e;
    di.dev = dev;
    cache[dev] = di;
  }
  *out = cache[dev];
  return 0;
}
// WSMG_FORCE_GENERIC=1 routes the reference shapes through the runtime-geometry kernel, WSMG_NO_TMA=1
// moves the map window with cp.async / st.global instead of TMA (both for tests and A/B profiling); read once.
// wsmg_debug_switches() overrides them at run time (tests compare the four builds inside one process).
struct Switches { bool generic, no_tma; };
static int g_switch_override = -1;          // -1: environment; else bit 0 = generic, bit 1 = no TMA
static Switches switches() {
  static const Switches env = [] {
    Switches s{};
    const char* f = getenv("WSMG_FORCE_GENERIC");
    const char* n = getenv("WSMG_NO_TMA");
    s.generic = f && f[0] == '1';
    s.no_tma = n && n[0] == '1';
    return s;
  }();
  const int o = g_switch_override;
  if (o < 0) return env;
  Switches s{(o & 1) != 0, (o & 2) != 0};
  return s;
}

template <int CE, int CG, int CHW, bool VEC, bool TMA, int FEAT = FEAT_NCHW>
static int launch_fused_t(const FusedParams& p, int grid, cudaStream_t s, int dev, bool pdl) {
  static int attr_set_for[64] = {0};          // dynamic shared memory opt-in, once per device and size
  if (attr_set_for[dev] < p.sp.total) {
    cudaError_t e = cudaFuncSetAttribute(k_fused<CE, CG, CHW, VEC, TMA, FEAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, p.sp.total);
    if (e != cudaSuccess) return (int)e;
    attr_set_for[dev] = p.sp.total;
  }
  if (!pdl) {
    k_fused<CE, CG, CHW, VEC, TMA, FEAT><<<grid, FUSED_NT, p.sp.total, s>>>(p);
    return (int)cudaGetLastError();
  }
  // Programmatic dependent launch behind k_cells: the CTAs are scheduled while k_cells' last blocks drain and run their
  // prologue (tables, key planes, pose) up to griddepcontrol.wait.
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(FUSED_NT); cfg.dynamicSmemBytes = (size_t)p.sp.total; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return (int)cudaLaunchKernelEx(&cfg, k_fused<CE, CG, CHW, VEC, TMA, FEAT>, p);
}

static int launch_fused(FusedParams p, const wsmg_dims* d, cudaStream_t s, bool pdl = false) {
  p.sp = geo_plan(d).sp;
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc != 0) return rc;
  if (p.sp.total > di.max_optin) return WSMG_E_SMEM;
  const bool vec = (p.g.C % 4) == 0;
  p.n_maps = d->n_maps;
  const int grid = p.bs * ((p.g.C + SLAB - 1) / SLAB);
  p.debug_skip = 0;
#if defined(WSMG_PHASE_SKIP)
  if (const char* dbg = getenv("WSMG_DEBUG_SKIP")) p.debug_skip = atoi(dbg);   // profiling build only, see WSMG_SKIP
#endif
  const Switches sw = switches();
  bool tma = vec && !p.stop_after_scatter && !sw.no_tma && p.sp.wwp <= p.g.G;   // box no wider than the map
  if (tma) {
    rc = encode_window_map(&p.tmap, p.gmap, d->n_maps, p.g, p.sp.wwp);
    if (rc != 0) return rc;
  }
  if (p.g.Cin != p.g.C && vec && tma && p.g.E == 100 && p.g.G == 240 && p.g.Hf * p.g.Wf == 224 * 224 && !sw.generic)
    return launch_fused_t<100, 240, 224 * 224, true, true, FEAT_POOL>(p, grid, s, di.dev, pdl);   // a wider producer at the reference's shapes
  if (p.g.Cin != p.g.C) {                                   // channel pool fused in the scatter: run-time geometry builds
    if (vec) return tma ? launch_fused_t<0, 0, 0, true, true, FEAT_POOL>(p, grid, s, di.dev, pdl) : launch_fused_t<0, 0, 0, true, false, FEAT_POOL>(p, grid, s, di.dev, pdl);
    return launch_fused_t<0, 0, 0, false, false, FEAT_POOL>(p, grid, s, di.dev, pdl);
  }
  if (p.g.feat_nhwc) {                                      // channels_last features (validated: C % 4 == 0, no pool)
    if (!aligned16(p.feat) && !p.stop_after_scatter && p.proj_in == nullptr) return WSMG_E_ALIGN;
    if (tma && p.g.E == 100 && p.g.G == 240 && p.g.Hf * p.g.Wf == 224 * 224 && !sw.generic)
      return launch_fused_t<100, 240, 224 * 224, true, true, FEAT_NHWC>(p, grid, s, di.dev, pdl);
    return tma ? launch_fused_t<0, 0, 0, true, true, FEAT_NHWC>(p, grid, s, di.dev, pdl) : launch_fused_t<0, 0, 0, true, false, FEAT_NHWC>(p, grid, s, di.dev, pdl);
  }
  const bool ref_geo = vec && p.g.E == 100 && p.g.G == 240 && !sw.generic;
  if (ref_geo && p.g.Hf * p.g.Wf == 224 * 224) {          // the reference's shapes (vlnce_task.yaml:11-18)
    return tma ? launch_fused_t<100, 240, 224 * 224, true, true>(p, grid, s, di.dev, pdl)
               : launch_fused_t<100, 240, 224 * 224, true, false>(p, grid, s, di.dev, pdl);
  }
  if (ref_geo && p.g.Hf * p.g.Wf == 256 * 256 && tma) {   // BASELINE.json's wording: features at the depth resolution
    return launch_fused_t<100, 240, 256 * 256, true, true>(p, grid, s, di.dev, pdl);
  }
  if (vec) return tma ? launch_fused_t<0, 0, 0, true, true>(p, grid, s, di.dev, pdl) : launch_fused_t<0, 0, 0, true, false>(p, grid, s, di.dev, pdl);
  // C % 4 != 0 (no 16-byte cells, no TMA) at the reference's grid sizes: compile-time geometry as well -- the run-time
  // build spends a third of its instructions on integer divisions by E, E + 2 and the tile counts
  if (p.g.E == 100 && p.g.G == 240 && !sw.generic) {
    if (p.g.Hf * p.g.Wf == 224 * 224) return launch_fused_t<100, 240, 224 * 224, false, false>(p, grid, s, di.dev, pdl);
    if (p.g.Hf * p.g.Wf == 256 * 256) return launch_fused_t<100, 240, 256 * 256, false, false>(p, grid, s, di.dev, pdl);
  }
  return launch_fused_t<0, 0, 0, false, false>(p, grid, s, di.dev, pdl);
}

}  // namespace wsmg

using namespace wsmg;

extern "C" {

int wsmg_abi_version(void) { return WSMG_ABI_VERSION; }

void wsmg_debug_switches(int force_generic, int no_tma) {
  g_switch_override = (force_generic < 0 || no_tma < 0) ? -1 : ((force_generic ? 1 : 0) | (no_tma ? 2 : 0));
}

const char* wsmg_error_string(int code) {
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  const char* s = error_string(code);
  return s ? s : "unknown wsmg error";
}

size_t wsmg_scratch_bytes(const wsmg_dims* d) {
  if (validate_dims(d) != WSMG_OK) return 0;
  return scratch_bytes(d);
}

size_t wsmg_scratch_flags_offset(const wsmg_dims* d) {
  if (validate_dims(d) != WSMG_OK) return 0;
  return scratch_codes_bytes(d);
}

int wsmg_base_coords_host(float* out_host, int32_t n) {
  if (out_host == nullptr) return WSMG_E_NULL;
  if (n <= 0) return WSMG_E_DIMS;
  base_coords_host(out_host, n);
  return WSMG_OK;
}

static int map_update_impl(const float* feat, const float* depth, const float* gps, const float* compass,
                           const float* mask, float* gmap, float* ego_out, const wsmg_opts* o, void* scratch,
                           size_t scratch_bytes_, const wsmg_dims* d, cudaStream_t s) {
  int rc = validate_dims(d);
  if (rc != WSMG_OK) return rc;
  if (!feat || !depth || !gps || !compass || !mask || !gmap || !ego_out || !scratch) return WSMG_E_NULL;
  if (!aligned16(feat) || !aligned16(gmap) || !aligned16(scratch)) return WSMG_E_ALIGN;
  if (scratch_bytes_ < scratch_bytes(d)) return WSMG_E_SCRATCH;
  const Geo& g = geo_plan(d).g;
  const ScratchView sv = scratch_view(scratch, d);
  PreArgs pre;
  pre.block_flags = sv.block_flags; pre.gmap = gmap; pre.mask = mask; pre.per_env = (size_t)d->G * d->G * d->C;
  pre.env_slots = o->env_slots; pre.row_bounds = sv.bounds; pre.env_trig = sv.env_trig; pre.compass = compass; pre.trig = o->trig;
  pre.n_maps = d->n_maps;
  rc = launch_cells(depth, sv.codes, nullptr, nullptr, nullptr, o->status, g, d->bs, s, &pre);
  if (rc) return rc;
  FusedParams p{};
  p.row_bounds = sv.bounds; p.env_trig = sv.env_trig; p.status = o->status;
  p.block_flags = sv.block_flags; p.flag_words = pre.cell_blocks + 1; p.env_flags_out = sv.flags;
  p.feat = feat; p.codes = sv.codes; p.env_flags = sv.flags; p.gps = gps; p.compass = compass; p.trig = o->trig;
  p.gmap = gmap; p.ego = ego_out; p.proj_out = nullptr; p.proj_in = nullptr;
  p.ego_half = (uint16_t*)o->ego_half; p.env_slots = o->env_slots;
  p.stop_after_scatter = 0; p.bs = d->bs; p.g = g;
  if (o->ev_before_fused) cudaEventRecord((cudaEvent_t)o->ev_before_fused, s);
  rc = launch_fused(p, d, s, /*pdl=*/o->ev_before_fused == nullptr);   // (an event between the two launches would serialise them anyway)
  if (o->ev_after_fused) cudaEventRecord((cudaEvent_t)o->ev_after_fused, s);
  return rc;
}

int wsmg_map_update(const float* feat, const float* depth, const float* gps, const float* compass,
                    const float* mask, float* gmap, float* ego_out, const float* trig, void* scratch,
                    size_t scratch_bytes_, const wsmg_dims* d, void* stream) {
  wsmg_opts o{};
  o.trig = trig;
  return map_update_impl(feat, depth, gps, compass, mask, gmap, ego_out, &o, scratch, scratch_bytes_, d, (cudaStream_t)stream);
}

int wsmg_map_update_ex(const float* feat, const float* depth, const float* gps, const float* compass,
                       const float* mask, float* gmap, float* ego_out, const wsmg_opts* o, void* scratch,
                       size_t scratch_bytes_, const wsmg_dims* d, void* stream) {
  wsmg_opts z{};
  if (o == nullptr) o = &z;
  return map_update_impl(feat, depth, gps, compass, mask, gmap, ego_out, o, scratch, scratch_bytes_, d, (cudaStream_t)stream);
}

int wsmg_unproject_index(const float* depth, int32_t* lin, uint8_t* invalid, const wsmg_dims* d, void* stream) {
  int rc = validate_dims(d);
  if (rc != WSMG_OK) return rc;
  if (!depth || !lin || !invalid) return WSMG_E_NULL;
  return launch_cells(depth, nullptr, lin, invalid, nullptr, nullptr, make_geo(d), d->bs, (cudaStream_t)stream);
}

int wsmg_scatter_max(const float* feat, const float* depth, float* proj_out, void* scratch, size_t scratch_bytes_,
                     const wsmg_dims* d, void* stream) {
  int rc = validate_dims(d);
  if (rc != WSMG_OK) return rc;
  if (!feat || !depth || !proj_out || !scratch) return WSMG_E_NULL;
  if (!aligned16(feat) || !aligned16(scratch)) return WSMG_E_ALIGN;
  if (scratch_bytes_ < scratch_bytes(d)) return WSMG_E_SCRATCH;
  cudaStream_t s = (cudaStream_t)stream;
  const Geo g = make_geo(d);
  const ScratchView sv = scratch_view(scratch, d);
  ResetArgs ra;
  ra.env_flags = sv.flags;
  rc = launch_reset(nullptr, nullptr, d, s, ra);      // only clears the env flags
  if (rc) return rc;
  rc = launch_cells(depth, sv.codes, nullptr, nullptr, sv.flags, nullptr, g, d->bs, s);
  if (rc) return rc;
  FusedParams p{};
  p.feat = feat; p.codes = sv.codes; p.env_flags = sv.flags; p.proj_out = proj_out; p.stop_after_scatter = 1; p.bs = d->bs; p.g = g;
  return launch_fused(p, d, s);
}

int wsmg_register_fuse_retrieve(const float* proj_in, const float* gps, const float* compass, const float* mask,
                                float* gmap, float* ego_out, const float* trig, void* scratch, size_t scratch_bytes_,
                                const wsmg_dims* d, void* stream) {
  int rc = validate_dims(d);
  if (rc != WSMG_OK) return rc;
  if (!proj_in || !gps || !compass || !mask || !gmap || !ego_out || !scratch) return WSMG_E_NULL;
  if (!aligned16(gmap) || !aligned16(scratch)) return WSMG_E_ALIGN;
  if (scratch_bytes_ < scratch_bytes(d)) return WSMG_E_SCRATCH;
  cudaStream_t s = (cudaStream_t)stream;
  const ScratchView sv = scratch_view(scratch, d);
  ResetArgs ra;
  ra.row_bounds = sv.bounds; ra.env_trig = sv.env_trig; ra.compass = compass; ra.trig = trig;
  rc = launch_reset(gmap, mask, d, s, ra);
  if (rc) return rc;
  FusedParams p{};
  p.row_bounds = sv.bounds; p.env_trig = sv.env_trig;
  p.proj_in = proj_in; p.gps = gps; p.compass = compass; p.trig = trig; p.gmap = gmap; p.ego = ego_out;
  p.bs = d->bs; p.g = make_geo(d);
  return launch_fused(p, d, s);
}

// ------------------------------------------------------------------ host-buffer entry
// Internal copy / compute lanes of the host-buffer entry: two non-blocking streams, a fork and two join events, per
// host thread and device, created on first use and kept for the life of the thread.
struct HostLanes { cudaStream_t st[2]; cudaEvent_t fork, join[2]; bool ready; };
static int host_lanes(HostLanes** out) {
  static thread_local HostLanes lanes[64];
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  if (dev < 0 || dev >= 64) return (int)cudaErrorInvalidDevice;
  HostLanes& l = lanes[dev];
  if (!l.ready) {
    if ((e = cudaEventCreateWithFlags(&l.fork, cudaEventDisableTiming)) != cudaSuccess) return (int)e;
    for (int i = 0; i < 2; ++i) {
      if ((e = cudaStreamCreateWithFlags(&l.st[i], cudaStreamNonBlocking)) != cudaSuccess) return (int)e;
      if ((e = cudaEventCreateWithFlags(&l.join[i], cudaEventDisableTiming)) != cudaSuccess) return (int)e;
    }
    l.ready = true;
  }
  *out = &l;
  return 0;
}

// All host -> device copies of one chunk as ONE cudaMemcpyBatchAsync (CUDA runtime >= 12.8, looked up at run time).
// Measured on PCIe Gen5 (profiles/r02_e2e_chunks.txt): every additional multi-megabyte cudaMemcpy(2D)Async leaves the
// H2D engine idle for ~17 us, which with one feature copy per env (row skipping) costs 9 % of a PCIe-bound step; the
// batch submits a chunk's large copies -- per env and plane the live rows, and the depth frames -- in one go
// (5.57 k -> 6.00 k frames/s at 128 envs per step, 94 % of the H2D bound).  WSMG_HOST_NO_BATCHCOPY=1 or an older runtime: one call per copy.
struct H2DList {
  std::vector<void*> dst, src;
  std::vector<size_t> size;
  void clear() { dst.clear(); src.clear(); size.clear(); }
  void add(void* d, const void* s, size_t bytes) {
    if (bytes == 0) return;
    dst.push_back(d); src.push_back(const_cast<void*>(s)); size.push_back(bytes);
  }
};
typedef cudaError_t (*memcpy_batch_fn)(void**, void**, size_t*, size_t, cudaMemcpyAttributes*, size_t*, size_t, size_t*, cudaStream_t);
static memcpy_batch_fn memcpy_batch() {
  static const memcpy_batch_fn fn = [] {
    const char* off = getenv("WSMG_HOST_NO_BATCHCOPY");
    if (off && off[0] == '1') return (memcpy_batch_fn) nullptr;
    // the CUDA runtime this library is bound to (it may sit in a local dlopen scope: look it up through one of its symbols)
    Dl_info info;
    void* h = nullptr;
    if (dladdr(reinterpret_cast<void*>(&cudaMemcpyAsync), &info) != 0 && info.dli_fname != nullptr)
      h = dlopen(info.dli_fname, RTLD_LAZY | RTLD_NOLOAD);
    void* sym = h != nullptr ? dlsym(h, "cudaMemcpyBatchAsync") : dlsym(RTLD_DEFAULT, "cudaMemcpyBatchAsync");
    return (memcpy_batch_fn)sym;
  }();
  return fn;
}
static cudaError_t submit_h2d(H2DList& l, cudaStream_t s) {
  if (l.dst.empty()) return cudaSuccess;
  static bool batch_refused = false;       // the runtime has the entry point but turned the call down once: stay with single copies
  memcpy_batch_fn fn = batch_refused ? nullptr : memcpy_batch();
  if (fn != nullptr) {
    cudaMemcpyAttributes at{};
    at.srcAccessOrder = cudaMemcpySrcAccessOrderStream;      // the host buffers are read in stream order, like cudaMemcpyAsync does
    size_t first = 0, fail = 0;
    const cudaError_t e = fn(l.dst.data(), l.src.data(), l.size.data(), l.dst.size(), &at, &first, 1, &fail, s);
    if (e == cudaSuccess) return e;
    if (e != cudaErrorNotSupported && e != cudaErrorInvalidValue) return e;
    cudaGetLastError();                    // argument-level refusal: nothing was enqueued, redo the chunk copy by copy
    batch_refused = true;
  }
  for (size_t i = 0; i < l.dst.size(); ++i) {
    cudaError_t e = cudaMemcpyAsync(l.dst[i], l.src[i], l.size[i], cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

// staging layout per chunk slot (2 slots): feat | depth | gps | compass | mask | ego | scratch
struct HostSlot { float *feat, *depth, *gps, *compass, *mask, *ego; void* scratch; size_t scratch_bytes; };

static size_t slot_bytes(const wsmg_dims* d, int chunk, HostSlot* out, unsigned char* base) {
  wsmg_dims dc = *d; dc.bs = chunk; dc.n_maps = chunk > d->n_maps ? chunk : d->n_maps;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  size_t o_feat = take((size_t)chunk * (d->C_in > 0 ? d->C_in : d->C) * d->Hf * d->Wf * 4);
  size_t o_depth = take((size_t)chunk * d->Hd * d->Wd * 4);
  size_t o_gps = take((size_t)chunk * 2 * 4);
  size_t o_comp = take((size_t)chunk * 4);
  size_t o_mask = take((size_t)chunk * 4);
  size_t o_ego = take((size_t)chunk * d->C * d->E * d->E * 4);
  size_t sb = scratch_bytes(&dc);
  size_t o_scr = take(sb);
  if (out && base) {
    out->feat = (float*)(base + o_feat); out->depth = (float*)(base + o_depth); out->gps = (float*)(base + o_gps);
    out->compass = (float*)(base + o_comp); out->mask = (float*)(base + o_mask); out->ego = (float*)(base + o_ego);
    out->scratch = base + o_scr; out->scratch_bytes = sb;
  }
  return off;
}

// Does sampled row i of one (host) depth frame hold a pixel that can write?  rgb_mapping.py:165-174 evaluated with the
// device's own arithmetic (wsmg_math.h).
static bool row_can_write(const Geo& g, const float* depth_env, const int* col_src, const float* col_xx, int i) {
  const int r = sample_index(g, i);
  const float yy = pinhole_yy(g, r);
  const float* row = depth_env + (size_t)r * g.Wd;
  // cheap conservative filter (one vector pass over the depth row): a pixel can only write if its height
  // Y = yy * Z lies in (-1.5, 0.1), i.e. if Z is below a per-row bound; rows whose smallest non-zero depth is
  // already beyond it (everything above the horizon of an indoor frame) are settled without the exact test
  float zmin = INFINITY;
  int c = 0;
#if defined(__SSE2__)
  {
    __m128 vmin = _mm_set1_ps(INFINITY);
    const __m128 inf = _mm_set1_ps(INFINITY), zero = _mm_setzero_ps();
    for (; c + 4 <= g.Wd; c += 4) {
      __m128 v = _mm_loadu_ps(row + c);
      const __m128 nz = _mm_cmpneq_ps(v, zero);                       // (NaN != 0 is true: a NaN stays and loses the min)
      v = _mm_or_ps(_mm_and_ps(nz, v), _mm_andnot_ps(nz, inf));
      vmin = _mm_min_ps(v, vmin);                                     // v < vmin ? v : vmin
    }
    float lane[4];
    _mm_storeu_ps(lane, vmin);
    for (int l = 0; l < 4; ++l) zmin = lane[l] < zmin ? lane[l] : zmin;
  }
#endif
  for (; c < g.Wd; ++c) {
    const float v = row[c] != 0.0f ? row[c] : INFINITY;
    zmin = v < zmin ? v : zmin;
  }
  const float bound = yy > 0.0f ? 0.1f / yy : (yy < 0.0f ? 1.5f / -yy : INFINITY);
  if (!(zmin * 10.0f < bound * 1.0001f)) return false;
  for (int j = 0; j < g.Wf; ++j) {
    int x, y;
    if (unproject_depth(g, row[col_src[j]], col_xx[j], yy, &x, &y)) return true;
  }
  return false;
}

// First and last such row (lo > hi: none), scanning inwards from both ends.  A superset would do -- the kernel only
// reads the features of pixels that pass exactly this test and land inside the fan.
static void live_rows_host(const Geo& g, const float* depth_env, const int* col_src, const float* col_xx, int* lo, int* hi) {
  *lo = g.Hf; *hi = -1;
  for (int i = 0; i < g.Hf; ++i)
    if (row_can_write(g, depth_env, col_src, col_xx, i)) { *lo = i; break; }
  if (*lo == g.Hf) return;
  for (int i = g.Hf - 1; i >= *lo; --i)
    if (row_can_write(g, depth_env, col_src, col_xx, i)) { *hi = i; break; }
}

size_t wsmg_host_staging_bytes(const wsmg_dims* d, int32_t chunk_envs) {
  if (validate_dims(d) != WSMG_OK || chunk_envs <= 0) return 0;
  int chunk = chunk_envs < d->bs ? chunk_envs : d->bs;
  return 2 * slot_bytes(d, chunk, nullptr, nullptr);
}

int wsmg_map_update_host(const float* feat_host, const float* depth_host, const float* gps_host,
                         const float* compass_host, const float* mask_host, float* gmap, float* ego_out_host,
                         void* staging, size_t staging_bytes, int32_t chunk_envs, const wsmg_dims* d, void* stream) {
  return wsmg_map_update_host_ex(feat_host, depth_host, gps_host, compass_host, mask_host, gmap, ego_out_host, staging,
                                 staging_bytes, chunk_envs, d, 0u, stream);
}

int wsmg_map_update_host_ex(const float* feat_host, const float* depth_host, const float* gps_host,
                            const float* compass_host, const float* mask_host, float* gmap, float* ego_out_host,
                            void* staging, size_t staging_bytes, int32_t chunk_envs, const wsmg_dims* d, uint32_t flags,
                            void* stream) {
  int rc = validate_dims(d);
  if (rc != WSMG_OK) return rc;
  if (!feat_host || !depth_host || !gps_host || !compass_host || !mask_host || !gmap || !ego_out_host || !staging)
    return WSMG_E_NULL;
  if (chunk_envs <= 0) return WSMG_E_DIMS;
  const int chunk = chunk_envs < d->bs ? chunk_envs : d->bs;
  if (staging_bytes < wsmg_host_staging_bytes(d, chunk) || !aligned16(staging)) return WSMG_E_SCRATCH;
  // zero-copy: the kernel reads the caller's pinned buffer through its device mapping
  const float* feat_mapped = nullptr;
  if (flags & WSMG_HOST_ZEROCOPY_FEATURES) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, feat_host) != cudaSuccess || at.type != cudaMemoryTypeHost || at.devicePointer == nullptr) {
      cudaGetLastError();
      return WSMG_E_HOSTMEM;
    }
    feat_mapped = static_cast<const float*>(at.devicePointer);
    if (!aligned16(feat_mapped)) return WSMG_E_ALIGN;
  }
  static_assert(CELLS_MAX_W <= 1024, "column tables live on the stack");
  int col_src[CELLS_MAX_W];
  float col_xx[CELLS_MAX_W];
  if (flags & WSMG_HOST_SKIP_DEAD_ROWS) {
    if (d->Wf > CELLS_MAX_W) return WSMG_E_DIMS;
    const Geo g = make_geo(d);
    for (int j = 0; j < d->Wf; ++j) { col_src[j] = sample_index(g, j); col_xx[j] = pinhole_xx(g, col_src[j]); }
  }
  cudaStream_t user = (cudaStream_t)stream;
  // Two internal streams ping-pong over the two staging slots so that the copies of chunk i+1 overlap the kernels of
  // chunk i; both are fenced against the caller's stream with events.  Streams and events are created once per host
  // thread and device and kept (a trainer calls this every step).
  HostLanes* lanes = nullptr;
  rc = host_lanes(&lanes);
  if (rc != 0) return rc;
  cudaStream_t* st = lanes->st;
  cudaError_t e;
  auto ck = [&](cudaError_t err) { if (err != cudaSuccess && rc == 0) rc = (int)err; };
  ck(cudaEventRecord(lanes->fork, user));
  ck(cudaStreamWaitEvent(st[0], lanes->fork, 0));
  ck(cudaStreamWaitEvent(st[1], lanes->fork, 0));
  const size_t one = slot_bytes(d, chunk, nullptr, nullptr);
  const size_t per_map = (size_t)d->G * d->G * d->C;
  // Chunk sizes ramp up from 2 envs to `chunk` and down again at the end: the first copy is short, so the kernels start
  // early, and so is the last one, so little is left to drain when the copy engine runs dry (measured, 128 envs, PCIe
  // Gen5: 5.38 k frames/s with uniform chunks of 16, 5.66 k with chunks of 2 -- the ramp gets the latter without its
  // sixty-four launches).
  int slot = 0, n = 0, ramp = chunk < 2 ? chunk : 2;
  for (int b0 = 0; b0 < d->bs && rc == 0; b0 += n, slot ^= 1, ramp = ramp * 2 < chunk ? ramp * 2 : chunk) {
    const int left = d->bs - b0;
    n = left <= 2 ? left : (ramp < (left + 1) / 2 ? ramp : (left + 1) / 2);
    if (n > chunk) n = chunk;
    HostSlot hs;
    slot_bytes(d, chunk, &hs, (unsigned char*)staging + slot * one);
    cudaStream_t s = st[slot];
    const size_t fe = (size_t)(d->C_in > 0 ? d->C_in : d->C) * d->Hf * d->Wf, de = (size_t)d->Hd * d->Wd, ee = (size_t)d->C * d->E * d->E;
    static thread_local H2DList h2d;
    h2d.clear();
    const bool batched = memcpy_batch() != nullptr;       // (a refusal at submit time falls back to one 1-D copy per plane)
    if (feat_mapped != nullptr) {
      // nothing to stage
    } else if (flags & WSMG_HOST_SKIP_DEAD_ROWS) {
      const Geo g = make_geo(d);
      const int planes = d->C_in > 0 ? d->C_in : d->C;
      const size_t pitch = (size_t)d->Hf * d->Wf * 4;
      for (int k = 0; k < n; ++k) {
        int lo, hi;
        live_rows_host(g, depth_host + (size_t)(b0 + k) * de, col_src, col_xx, &lo, &hi);
        if (lo > hi) continue;                                   // nothing in this frame can write
        const size_t first = (size_t)lo * d->Wf;
        if (d->feat_nhwc) {                                      // rows lo..hi of an NHWC frame are one contiguous span
          const size_t off = first * planes, cnt = (size_t)(hi - lo + 1) * d->Wf * planes;
          h2d.add(hs.feat + (size_t)k * fe + off, feat_host + (size_t)(b0 + k) * fe + off, cnt * 4);
          continue;
        }
        if (batched) {                                           // the live rows of every plane: `planes` spans per env
          for (int pl = 0; pl < planes; ++pl)
            h2d.add(hs.feat + (size_t)k * fe + (size_t)pl * d->Hf * d->Wf + first,
                    feat_host + (size_t)(b0 + k) * fe + (size_t)pl * d->Hf * d->Wf + first, (size_t)(hi - lo + 1) * d->Wf * 4);
        } else {
          ck(cudaMemcpy2DAsync(hs.feat + (size_t)k * fe + first, pitch, feat_host + (size_t)(b0 + k) * fe + first, pitch,
                               (size_t)(hi - lo + 1) * d->Wf * 4, planes, cudaMemcpyHostToDevice, s));
        }
      }
    } else {
      h2d.add(hs.feat, feat_host + b0 * fe, n * fe * 4);
    }
    h2d.add(hs.depth, depth_host + b0 * de, n * de * 4);
    ck(submit_h2d(h2d, s));
    // (the few bytes of pose stay ordinary copies: inside the batch they cost 20 % of the step, measured)
    ck(cudaMemcpyAsync(hs.gps, gps_host + b0 * 2, (size_t)n * 2 * 4, cudaMemcpyHostToDevice, s));
    ck(cudaMemcpyAsync(hs.compass, compass_host + b0, (size_t)n * 4, cudaMemcpyHostToDevice, s));
    ck(cudaMemcpyAsync(hs.mask, mask_host + b0, (size_t)n * 4, cudaMemcpyHostToDevice, s));
    if (rc != 0) break;
    wsmg_dims dc = *d; dc.bs = n; dc.n_maps = n;
    rc = wsmg_map_update(feat_mapped != nullptr ? feat_mapped + b0 * fe : hs.feat, hs.depth, hs.gps, hs.compass, hs.mask, gmap + b0 * per_map, hs.ego, nullptr,
                         hs.scratch, hs.scratch_bytes, &dc, s);
    if (rc == 0) ck(cudaMemcpyAsync(ego_out_host + b0 * ee, hs.ego, n * ee * 4, cudaMemcpyDeviceToHost, s));
  }
  // always rejoin the caller's stream, also after an error: nothing may outlive the call unordered
  for (int i = 0; i < 2; ++i) {
    e = cudaEventRecord(lanes->join[i], st[i]);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(user, lanes->join[i], 0);
    ck(e);
  }
  if (rc == 0) rc = (int)cudaGetLastError();
  return rc;
}

int wsmg_semantic_crop(const float* maps, const float* pose, const float* trig, const int32_t* map_index, int64_t* out,
                       int32_t bs, int32_t n_maps, int32_t S, int32_t half, int32_t origin, void* stream) {
  if (!maps || !pose || !out) return WSMG_E_NULL;
  if (bs <= 0 || n_maps <= 0 || S <= 0 || S > 32768 || half <= 0 || half > 16384) return WSMG_E_DIMS;
  if (map_index == nullptr && n_maps < bs) return WSMG_E_BATCH;
  const int cells = 4 * half * half;
  dim3 grid((cells + 255) / 256, bs);
  k_semcrop<<<grid, 256, 0, (cudaStream_t)stream>>>(maps, pose, trig, map_index, reinterpret_cast<long long*>(out), n_maps, S,
                                                     half, origin);
  return (int)cudaGetLastError();
}

int wsmg_host_live_rows(const float* depth_host, const wsmg_dims* d, int32_t* row_lo, int32_t* row_hi) {
  int rc = validate_dims(d);
  if (rc != WSMG_OK) return rc;
  if (!depth_host || !row_lo || !row_hi) return WSMG_E_NULL;
  if (d->Wf > CELLS_MAX_W) return WSMG_E_DIMS;
  const Geo g = make_geo(d);
  int col_src[CELLS_MAX_W];
  float col_xx[CELLS_MAX_W];
  for (int j = 0; j < d->Wf; ++j) { col_src[j] = sample_index(g, j); col_xx[j] = pinhole_xx(g, col_src[j]); }
  for (int b = 0; b < d->bs; ++b) {
    int lo, hi;
    live_rows_host(g, depth_host + (size_t)b * d->Hd * d->Wd, col_src, col_xx, &lo, &hi);
    row_lo[b] = lo; row_hi[b] = hi;
  }
  return WSMG_OK;
}

}  // extern "C"
