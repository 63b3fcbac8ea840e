/*
 * wsmg.h -- C ABI of libwsmg.so: the WS-MGMap per-step map update on B200 (sm_100a).
 *
 * The reference has no FFI layer: its boundary for this path is the Python
 * module vlnce_baselines/common/rgb_mapping.py::RGBMapping.
 * These entry points are what that module's forward() binds to (through
 * ctypes, see ws-mgmap_b200/_lib.py and INTEGRATION.md); each comment cites the
 * reference lines the call replaces.  Plain pointers and sizes only; all
 * `const float*` / `float*` arguments are DEVICE pointers unless a name ends
 * in `_host`.  Every call is stream-ordered on `stream` (a cudaStream_t passed
 * as void*), never synchronises the device, keeps no global mutable state and
 * owns no memory: scratch is caller-allocated (wsmg_scratch_bytes).
 *
 * Return value: 0 = success; <0 = argument error (WSMG_E_*); >0 = cudaError_t
 * of the failed runtime call / launch.  wsmg_error_string() decodes both.
 */
#ifndef WSMG_H_
#define WSMG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WSMG_ABI_VERSION 3

enum {
  WSMG_OK = 0,
  WSMG_E_NULL = -1,        /* a required pointer is NULL                          */
  WSMG_E_DIMS = -2,        /* non-positive / inconsistent dimension               */
  WSMG_E_EGO_GT_GLOBAL = -3, /* egocentric map larger than the global map           */
  WSMG_E_CHANNELS = -4,    /* channel count not supported                         */
  WSMG_E_SMEM = -5,        /* geometry needs more shared memory than one SM has   */
  WSMG_E_ALIGN = -6,       /* pointer not 16-byte aligned / Hf*Wf not multiple of 4 */
  WSMG_E_SCRATCH = -7,     /* scratch buffer too small                            */
  WSMG_E_BATCH = -8,       /* bs larger than the map tensor's leading dimension   */
  WSMG_E_HOSTMEM = -9      /* zero-copy asked for, but the host buffer is not device-mapped pinned memory */
};

/* Geometry of one call.  Mirrors Mapping.__init__ (rgb_mapping.py:12-30) and the
 * tensor shapes of RGBMapping.forward (rgb_mapping.py:79-90). */
typedef struct wsmg_dims {
  int32_t bs;      /* frames in this call (envs)                                  */
  int32_t n_maps;  /* leading dim of the caller's full_global_map tensor (>= bs)   */
  int32_t C;       /* map_depth: channels of features, global map and ego map      */
  int32_t Hf, Wf;  /* feature frame (UNet proj_feat), NCHW                         */
  int32_t Hd, Wd;  /* depth frame                                                  */
  int32_t E;       /* egocentric_map_size (100)                                    */
  int32_t G;       /* global_map_size (240)                                        */
  double resolution; /* metres per cell (0.12)                                      */
  int32_t C_in;    /* channels of `feat` (0 = C).  C_in != C: the channel re-binning of
                      RGBMapping.forward (adaptive_max_pool1d over channels, rgb_mapping.py:81-84)
                      is applied inside the scatter: output channel k = max over input channels
                      [floor(k*C_in/C), ceil((k+1)*C_in/C)).                       */
  int32_t feat_nhwc; /* 0: `feat` is [bs,C,Hf,Wf] (NCHW, what unet_encoder.py:103-111 returns by default);
                      1: `feat` is [bs,Hf,Wf,C] in memory -- a torch channels_last tensor, the layout cuDNN's NHWC
                      convolutions produce -- and is consumed as it is, no permute copy.  Needs C % 4 == 0 and
                      C_in == 0 / C (no channel re-binning).                        */
} wsmg_dims;

int wsmg_abi_version(void);
const char* wsmg_error_string(int code);

/* Test / profiling hook, the library's only process-wide state: which build of k_fused serves the reference shapes.
 * force_generic = 1: the run-time-geometry kernel instead of the compile-time one; no_tma = 1: the map window moves
 * with cp.async / st.global instead of TMA.  Results are bit-identical (tests/test_gpu_parity.py compares all four).
 * Negative arguments return to the default: the environment variables WSMG_FORCE_GENERIC / WSMG_NO_TMA, read once. */
void wsmg_debug_switches(int force_generic, int no_tma);

/* Bytes of device scratch wsmg_map_update needs for `d` (0.10 MB per env at the reference shapes): packed cell codes
 * (uint16 per sampled pixel), per-env and per-block flag words, the first rotation's column bounds and both rotations'
 * cos / sin per env.  Contents need no initialisation and carry nothing from one call to the next. */
size_t wsmg_scratch_bytes(const wsmg_dims* d);

/* Per-env status words the update leaves in scratch (uint32 each, at byte offset wsmg_scratch_flags_offset(d)):
 *   WSMG_FLAG_INVALID_PIXEL  some pixel of the frame does not write (normal: sky, holes, out of range)
 *   WSMG_FLAG_OUTSIDE_FAN    a *valid* pixel fell outside the packed fan the scatter keeps in shared memory.
 *                            Only depth < 0 can do that (Habitat depth is in [0,1]); such pixels are DROPPED,
 *                            which the reference would not do -- callers that cannot rule them out should check
 *                            this bit (the Python module raises when `strict_inputs` is set).
 *   WSMG_FLAG_BAD_SLOT       wsmg_opts.env_slots[b] is not a row of the map tensor (>= n_maps or negative): the
 *                            frame was SKIPPED (map untouched, ego row not written).
 * wsmg_opts.status (optional) receives the same two problem bits without a device synchronisation. */
#define WSMG_FLAG_INVALID_PIXEL 1u
#define WSMG_FLAG_OUTSIDE_FAN 2u
#define WSMG_FLAG_BAD_SLOT 4u
size_t wsmg_scratch_flags_offset(const wsmg_dims* d);

/* Whole step: Mapping.project_feat_to_map (rgb_mapping.py:32-72) as called by
 * RGBMapping.forward (rgb_mapping.py:85).
 *   feat     [bs,C_in,Hf,Wf] fp32 NCHW       (rgb_features as the UNet emits them; channel pool of :81-84 fused),
 *            or [bs,Hf,Wf,C] when wsmg_dims.feat_nhwc is set (channels_last producer)
 *   depth    [bs,Hd,Wd,1] fp32 in [0,1]      (observations['depth']; the x10 of :37 is applied inside)
 *   gps      [bs,2], compass [bs,1], mask [bs,1] fp32
 *   gmap     [n_maps,G,G,C] fp32 NHWC        (self.full_global_map; rows [:bs] updated in place, :35,:56)
 *   ego_out  [bs,C,E,E] fp32 NCHW            (final_retrieval, :70)
 *   trig     optional [bs,4] = cos(-compass), sin(-compass), cos(compass), sin(compass) computed by
 *            the caller (parity tests pass the CPU reference's values); NULL = sinf/cosf on device.
 * Preconditions under which the result equals the reference's (all hold for the reference's own data flow):
 *   - gmap >= 0 everywhere (it starts at zero and only ever takes maxima of post-ReLU features or zeros): the update
 *     reads and writes only the (E+2)^2 window the ego patch can reach, whereas rgb_mapping.py:55-56 takes
 *     max(map, translated) over the whole G x G map, which would clamp a negative entry anywhere to 0;
 *   - features, map, gps and compass finite (torch.max propagates NaN; fmaxf does not), depth >= 0 (see
 *     WSMG_FLAG_OUTSIDE_FAN above);
 *   - mask in {0, 1} as the trainers produce it (other values scale the map as the reference does, :35).
 */
int wsmg_map_update(const float* feat, const float* depth, const float* gps, const float* compass,
                    const float* mask, float* gmap, float* ego_out, const float* trig,
                    void* scratch, size_t scratch_bytes, const wsmg_dims* d, void* stream);

/* Optional extras of wsmg_map_update_ex (all fields may be NULL / 0).
 *   trig        see wsmg_map_update.
 *   ego_half    [bs,C,E,E] fp16 (IEEE binary16, round-to-nearest-even) copy of ego_out, written by the same
 *               kernel: what the rollout store keeps (reference common_trainer.py:519-520 casts the fp32 map
 *               with numpy on the CPU after the forward hook's o.cpu(), dagger_trainer.py:303-306).
 *   env_slots   [bs] int32: frame b reads / updates map row env_slots[b] (< n_maps) instead of row b, so
 *               pausing finished envs is an index-table edit instead of the reference's
 *               full_global_map[state_index] re-materialisation (common_trainer.py:171-172,454-476).
 *               Slots must be distinct (two frames on one row race); a slot outside [0, n_maps) skips its frame
 *               and raises WSMG_FLAG_BAD_SLOT.
 *   ev_before_fused / ev_after_fused   cudaEvent_t recorded around the k_fused launch (profiling; bench.py times
 *               the dominant kernel with them).
 *   status      device-ACCESSIBLE uint32[2] the kernels raise without any synchronisation -- meant to be pinned,
 *               mapped host memory (cudaHostAlloc / torch pin_memory) that the caller polls from the CPU at its
 *               next call: status[0] != 0: a valid pixel fell outside the fan and was dropped (WSMG_FLAG_OUTSIDE_FAN);
 *               status[1] != 0: an env slot was out of range (WSMG_FLAG_BAD_SLOT).  Sticky until the caller clears it. */
typedef struct wsmg_opts {
  const float* trig;
  void* ego_half;
  const int32_t* env_slots;
  void* ev_before_fused;
  void* ev_after_fused;
  uint32_t* status;
} wsmg_opts;

int wsmg_map_update_ex(const float* feat, const float* depth, const float* gps, const float* compass,
                       const float* mask, float* gmap, float* ego_out, const wsmg_opts* opts,
                       void* scratch, size_t scratch_bytes, const wsmg_dims* d, void* stream);

/* Stage: ComputeSpatialLocs.forward + the index half of ProjectToGroundPlane.forward
 * (rgb_mapping.py:153-176, 188-217).  Outputs per sampled pixel of the Hf x Wf frame:
 *   lin      [bs,Hf,Wf] int32  y_gp*E + x_gp, invalid writes forced to 0 (:207-208,:216)
 *   invalid  [bs,Hf,Wf] uint8  1 = invalid_writes (:204)
 */
int wsmg_unproject_index(const float* depth, int32_t* lin, uint8_t* invalid,
                         const wsmg_dims* d, void* stream);

/* Stage: scatter half of ProjectToGroundPlane.forward (rgb_mapping.py:210-232):
 *   proj_out [bs,C,E,E] fp32 NCHW = proj_feats (before RotateTensor). */
int wsmg_scatter_max(const float* feat, const float* depth, float* proj_out,
                     void* scratch, size_t scratch_bytes, const wsmg_dims* d, void* stream);

/* Stage: everything after the projection (rgb_mapping.py:35, 37(rotate via :267), 40-70):
 * rotate(-compass), paste, translate, mask + max-fuse into gmap, translate back, crop, rotate(+compass).
 *   proj_in  [bs,C,E,E] fp32 NCHW = proj_feats, zero outside the fan a depth >= 0 pixel can reach (what
 *            wsmg_scatter_max produces; other cells are ignored).  scratch as for wsmg_map_update. */
int wsmg_register_fuse_retrieve(const float* proj_in, const float* gps, const float* compass,
                                const float* mask, float* gmap, float* ego_out, const float* trig,
                                void* scratch, size_t scratch_bytes, const wsmg_dims* d, void* stream);

/* Host helper: ATen's affine_grid base coordinates, linspace(-1,1,n)*(n-1)/n (align_corners=False),
 * as torch-CPU produces them.  Used by tests to pin the in-kernel tables. */
int wsmg_base_coords_host(float* out_host, int32_t n);

/* End-to-end entry with HOST buffers (the call a trainer without device-resident
 * observations makes): copies feat/depth/gps/compass/mask host->device, runs
 * wsmg_map_update, copies ego_out device->host, all on `stream`, in env chunks so
 * copies overlap compute.  gmap stays device-resident (it is state, rgb_mapping.py:29).
 * `staging` is caller-allocated device memory of wsmg_host_staging_bytes(d, chunk) bytes. */
size_t wsmg_host_staging_bytes(const wsmg_dims* d, int32_t chunk_envs);
int wsmg_map_update_host(const float* feat_host, const float* depth_host, const float* gps_host,
                         const float* compass_host, const float* mask_host, float* gmap,
                         float* ego_out_host, void* staging, size_t staging_bytes,
                         int32_t chunk_envs, const wsmg_dims* d, void* stream);


/* wsmg_map_update_host with options.  WSMG_HOST_ZEROCOPY_FEATURES: `feat_host` must be page-locked, device-mapped
 * host memory (cudaHostAlloc / cudaHostRegister, torch `pin_memory()`); the scatter then pulls the features over
 * the bus itself and the 4-pixel groups that cannot write (typically 45-50 % of a frame) never cross it.  Replaces
 * the H2D half of batch_obs (common_trainer / dagger_trainer) for the feature tensor.
 * WSMG_HOST_SKIP_DEAD_ROWS: before the copies are queued, the host runs the unprojection test of rgb_mapping.py:165-174
 * on the depth frame (same arithmetic as the device) and copies, per env, only the feature rows between the first
 * and the last row that holds a pixel which can write; the rows above the horizon of an indoor frame -- typically
 * half of the 12.8 MB -- never cross the bus.  Results are identical.  Everything else as above. */
enum { WSMG_HOST_ZEROCOPY_FEATURES = 1, WSMG_HOST_SKIP_DEAD_ROWS = 2 };
int wsmg_map_update_host_ex(const float* feat_host, const float* depth_host, const float* gps_host,
                            const float* compass_host, const float* mask_host, float* gmap,
                            float* ego_out_host, void* staging, size_t staging_bytes,
                            int32_t chunk_envs, const wsmg_dims* d, uint32_t flags, void* stream);

/* Ground-truth semantic map sensor, batched.  Replaces the per-env, per-step CPU work of
 * GtSemanticMapSensor.get_observation, habitat_extensions/sensors.py:403-410:
 *   grid_sample(grid_sample(map, trans_grid, nearest), rot_grid, nearest), pad by `half`, crop
 *   [origin-half, origin+half)^2, .long()
 * maps   [n_maps,S,S] fp32 class ids (the episode's map, already rotated by the start heading, sensors.py:390-392)
 * pose   [bs,3] fp32 = ((grid_y-S/2)/(S/2), (grid_x-S/2)/(S/2), -heading)        (sensors.py:397-401)
 * trig   optional [bs,2] fp32 = (cos, sin) of pose[:,2] as the caller's host evaluated them (bit-exact parity with
 *        the reference's CPU result); NULL: cosf/sinf on the device
 * map_index optional [bs] int32: which map each frame reads (default b)
 * out    [bs,2*half,2*half] int64.  All pointers device memory; stream-ordered. */
int wsmg_semantic_crop(const float* maps, const float* pose, const float* trig, const int32_t* map_index,
                       int64_t* out, int32_t bs, int32_t n_maps, int32_t S, int32_t half, int32_t origin,
                       void* stream);

/* The host-side test WSMG_HOST_SKIP_DEAD_ROWS applies, exposed for callers and for byte accounting: for every frame of
 * the HOST depth tensor [bs,Hd,Wd,1], the first and the last sampled feature row holding a pixel that can write
 * (row_lo[b] > row_hi[b]: none).  Pure host code, no CUDA call. */
int wsmg_host_live_rows(const float* depth_host, const wsmg_dims* d, int32_t* row_lo, int32_t* row_hi);

#ifdef __cplusplus
}
#endif
#endif /* WSMG_H_ */
