#!/usr/bin/env python
"""Benchmark of the map update (BASELINE.json metric: map-update frames/sec at 1/2/4/8 B200 +
achieved HBM GB/s vs peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--workload sweep1024|env8]

Workload `sweep1024` is BASELINE.json configs[3] as written, the configuration the metric is quoted on: 1024 envs at
the reference shapes (C=64 features 224x224, depth 256x256, ego 100, global 240), SHARDED 1024/N per GPU (strong
scaling, the way dagger_trainer.py:259-262 splits a fixed dataset across ranks), no cross-GPU traffic in the loop, one
all_gather of the per-rank stats at the end.  At N > 1 the same run also times the weak-scaling variant (1024 envs per
GPU, every rank owning its envs like a rank of the reference owns its NUM_PROCESSES envs) and reports it as `weak`.
A step is one map update of every env (1 frame = 1 env x 1 step).  Inputs (1.6 GB of features per 128 envs) are far
larger than the 126 MB L2, so no L2 flush is needed between iterations.  `env8` is configs[1] (8 envs on one GPU,
L2 flushed between steps); the default run reports it as `small_batch`.

One JSON line on stdout (rank 0).  --impl reference times the reference's own rgb_mapping.py (the unmodified file,
staged under baseline/_ref/ by __graft_entry__.build(); the oracle port when it is absent) on the host cores.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SHAPES = {
    "real": dict(C=64, Hf=224, Wf=224, Hd=256, Wd=256, E=100, G=240, resolution=0.12),       # the reference's config
    "b256": dict(C=64, Hf=256, Wf=256, Hd=256, Wd=256, E=100, G=240, resolution=0.12),       # BASELINE.json wording
    "b256c27": dict(C=27, Hf=256, Wf=256, Hd=256, Wd=256, E=100, G=240, resolution=0.12),    # "reference class count"
}
SHAPE = dict(SHAPES["real"])
DEPTH_KINDS = ("uniform", "near", "room2", "room4")
CPU_SAMPLE_ENVS = 8          # envs per step of every CPU / stock-PyTorch comparator (the reference's own batch size)


def algorithmic_bytes_per_frame(s=SHAPE):
    """read every feature and depth element once, read-modify-write only the map window
    the ego patch can touch, write the ego map once, plus the pose scalars."""
    return 4 * (s["C"] * s["Hf"] * s["Wf"] + s["Hd"] * s["Wd"] + 2 * s["E"] ** 2 * s["C"] + s["E"] ** 2 * s["C"]) + 20


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic_per_env():
    """DRAM bytes per env of k_fused from the committed `ncu --set full` capture (profiles/ncu_traffic.json,
    written by scripts/ncu_summary.sh); (None, None) when no capture is committed."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        j = json.load(open(p))
        return float(j["k_fused_dram_bytes_per_env"]), j.get("source", "profiles/ncu_traffic.json")
    except Exception:
        return None, None


class ClockSampler:
    """SM clock / throttle reasons sampled with NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.01)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------- comparators
def _comparator_inputs(n_envs, seed):
    import torch
    import wsmgmap_b200  # noqa: F401
    from wsmgmap_b200.synth import RandomWalk, make_depth, make_features
    s = SHAPE
    gen = torch.Generator().manual_seed(seed)
    feat = make_features(n_envs, s["C"], s["Hf"], s["Wf"], gen)
    depth = torch.cat([make_depth(DEPTH_KINDS[b % 4], 1, s["Hd"], s["Wd"], gen) for b in range(n_envs)], 0)
    return feat, depth, RandomWalk(n_envs, seed=seed)


def reference_stepper(n_envs, device):
    """(step(feat, depth, gps, compass, masks) -> ego, kind): the reference's own RGBMapping.forward from the unmodified
    file when it is available (/root/reference here, baseline/_ref/ on the GPU box), else the oracle port of it."""
    from oracle import reference_loader as rl
    s = SHAPE
    if rl.reference_available() and s["C"] == 64:
        m, _ = rl.make_reference_mapper(n_envs, device=device, map_depth=s["C"], egocentric_map_size=s["E"],
                                        global_map_size=s["G"], resolution=s["resolution"])

        def step(feat, depth, gps, compass, masks):
            return m(feat, dict(depth=depth, gps=gps, compass=compass), masks)      # rgb_mapping.py:79-90
        return step, "reference"
    from oracle.mapping_oracle import OracleMapper
    orc = OracleMapper(n_envs, s["C"], device=device)
    return (lambda feat, depth, gps, compass, masks: orc.step(feat, depth, gps, compass, masks)), "port"


def cpu_reference_fps(n_envs, steps, warmup, seed=0):
    """The reference's PyTorch path on the host cores, all threads, `n_envs` envs per step of the same synthetic
    workload.  Returns (frames/s, cores, seconds per step, kind)."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    feat, depth, walk = _comparator_inputs(n_envs, seed)
    step, kind = reference_stepper(n_envs, "cpu")
    times = []
    with torch.no_grad():
        for t in range(warmup + steps):
            gps, compass, masks = walk.step()
            t0 = time.perf_counter()
            step(feat, depth, gps, compass, masks)
            dt = time.perf_counter() - t0
            if t >= warmup:
                times.append(dt)
    total = sum(times)
    return n_envs * len(times) / total, cores, total / len(times), kind


def run_reference(args, rank, world):
    if rank != 0:
        return
    import warnings
    warnings.filterwarnings("ignore", message=".*align_corners.*")
    n = CPU_SAMPLE_ENVS
    fps, cores, sec, kind = cpu_reference_fps(n, args.steps, max(args.warmup, 1))
    what = "the unmodified reference file rgb_mapping.py (RGBMapping.forward)" if kind == "reference" else "oracle port of rgb_mapping.py"
    sample = (f"{n} envs per step x {args.steps} steps (the reference's own batch size; frames/s does not depend on the env "
              f"count beyond it), {what}, torch CPU, {cores} threads, shapes and depth mix of the native arm")
    cfg = workload_config(args, world)
    cfg["reference_sample"] = sample
    line = {
        "metric": "map-update frames/sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": cfg["scaling"],
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
        "config": cfg,
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(local_rank):
    """Run this rank on the CPUs of the NUMA node its GPU hangs off, so that the pinned host buffers of the e2e leg are
    first-touched in that node's memory (what a production launcher does with numactl).  Returns (node, why): on a
    single-node host (the driver's box reports numa_node = -1 for every GPU) there is nothing to bind."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        bdf = pynvml.nvmlDeviceGetPciInfo(h).busId
        bdf = (bdf.decode() if isinstance(bdf, bytes) else bdf).lower()
        if len(bdf.split(":")[0]) == 8:
            bdf = bdf[4:]
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None, f"sysfs reports numa_node {node} for {bdf}: single-node host, nothing to bind"
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.extend(range(int(a), int(b or a) + 1))
        allowed = set(os.sched_getaffinity(0)) & set(cpus)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node, f"bound to the {len(allowed)} CPUs of node {node}"
        return None, f"node {node} has no CPU this process may run on"
    except Exception as e:                                       # best effort
        return None, f"not determined ({type(e).__name__})"


def workload_config(args, world):
    """sweep1024: 1024 envs in total, sharded 1024 / N per GPU (strong scaling, BASELINE.json configs[3]); --scaling
    weak gives every GPU the full 1024.  env8: 8 envs per GPU."""
    base = 1024 if args.workload == "sweep1024" else 8
    if args.envs:
        base = args.envs
    scaling = args.scaling or ("strong" if args.workload == "sweep1024" else "weak")
    if scaling == "weak":
        per_gpu, total = base, base * world
    else:
        per_gpu, total = base // world, base
    return {"workload": f"{args.workload}: {total} envs in total, {per_gpu} per GPU ({scaling} scaling), env-sharded, "
                        f"C={SHAPE['C']} feat {SHAPE['Hf']}x{SHAPE['Wf']} depth {SHAPE['Hd']}x{SHAPE['Wd']} ego 100 global 240 fp32, depth kinds mixed "
                        f"{'/'.join(DEPTH_KINDS)}, random-walk poses, a new frame per env per step, masks=1 after the first step",
            "envs_total": total, "envs_per_gpu": per_gpu, "scaling": scaling, "feat_layout": getattr(args, "feat_layout", "nchw"),
            "l2": "inputs larger than L2 (no flush)" if per_gpu >= 64 else "L2 flushed between steps",
            "bytes_per_frame_algorithmic": algorithmic_bytes_per_frame()}


# ------------------------------------------------------------------------------------------------- native arm
class Resident:
    """`n` envs of the workload resident in HBM and the step that updates them."""

    def __init__(self, n, T, dev, rank, lib, nhwc=False):
        import torch
        import wsmgmap_b200  # noqa: F401
        from wsmgmap_b200 import ops
        from wsmgmap_b200.synth import RandomWalk, make_depth
        s = SHAPE
        self.n, self.dev, self.lib = n, dev, lib
        gen = torch.Generator(device=dev).manual_seed(1234 + rank)
        # Every env sees a NEW frame every step, as in a rollout: n + T frames are resident and step t reads the
        # window [t, t+n).  (Re-feeding the same frame would leave the max-fused map unchanged after a few steps,
        # and a map update that changes nothing also writes nothing.)
        self.nhwc = nhwc
        if nhwc:      # channels_last producer: [frames, Hf, Wf, C] in memory (wsmg_dims.feat_nhwc)
            self.feat_all = torch.rand(n + T, s["Hf"], s["Wf"], s["C"], generator=gen, device=dev)
        else:
            self.feat_all = torch.rand(n + T, s["C"], s["Hf"], s["Wf"], generator=gen, device=dev)
        cgen = torch.Generator().manual_seed(99 + rank)
        kinds = [make_depth(k, 8, s["Hd"], s["Wd"], cgen) for k in DEPTH_KINDS]
        self.depth_all = torch.stack([kinds[b % 4][(b // 4) % 8] for b in range(n + T)], 0).to(dev).contiguous()
        walk = RandomWalk(n, seed=7 + rank)
        poses = [walk.step() for _ in range(T)]
        self.gps = torch.stack([p[0] for p in poses]).to(dev)
        self.compass = torch.stack([p[1] for p in poses]).to(dev)
        self.masks = torch.stack([p[2] for p in poses]).to(dev)
        self.gmap = torch.zeros(n, s["G"], s["G"], s["C"], device=dev)
        self.ego = torch.empty(n, s["C"], s["E"], s["E"], device=dev)
        self.d = ops.dims_for((n, s["C"], s["Hf"], s["Wf"]), (n, s["Hd"], s["Wd"], 1), n, s["E"], s["G"], s["resolution"], feat_nhwc=nhwc)
        self.scratch = ops.alloc_scratch(self.d, dev)
        self.stream = torch.cuda.current_stream(dev)

    def step(self, t, ev0=None, ev1=None):
        from wsmgmap_b200 import _lib
        P = lambda x: ctypes.c_void_p(x.data_ptr())  # noqa: E731
        n = self.n
        opts = _lib.WsmgOpts(None, None, None, None if ev0 is None else ctypes.c_void_p(ev0.cuda_event),
                             None if ev1 is None else ctypes.c_void_p(ev1.cuda_event), None)
        rc = self.lib.wsmg_map_update_ex(P(self.feat_all[t:t + n]), P(self.depth_all[t:t + n]), P(self.gps[t]), P(self.compass[t]),
                                         P(self.masks[t]), P(self.gmap), P(self.ego), ctypes.byref(opts), P(self.scratch),
                                         self.scratch.numel(), ctypes.byref(self.d), ctypes.c_void_p(self.stream.cuda_stream))
        _lib.check(rc, "wsmg_map_update_ex")


def time_resident(res, K, W, barrier, flush_buf=None):
    """W warm-up steps, then K timed steps (CUDA events on the launching stream), then K more with events around
    k_fused.  Returns (ms for the K steps, mean k_fused ms, checksum)."""
    import torch
    stream, dev = res.stream, res.dev
    for t in range(W):
        res.step(t)
    barrier()
    if flush_buf is not None:
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        for k in range(K):
            flush_buf.fill_(k & 0xFF)
            evs[k][0].record(stream)
            res.step(W + k)
            evs[k][1].record(stream)
        barrier()
        elapsed_ms = sum(a.elapsed_time(b) for a, b in evs)
    else:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for k in range(K):
            res.step(W + k)
        e1.record(stream)
        barrier()
        elapsed_ms = e0.elapsed_time(e1)
    checksum = float(res.ego.double().sum().item()) + float(res.gmap[0].double().sum().item())
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for a, b in kev:
        a.record(stream)
        b.record(stream)       # materialise the handles
    torch.cuda.synchronize(dev)
    for k in range(K):
        if flush_buf is not None:
            flush_buf.fill_(k & 0xFF)
        res.step(W + K + k, kev[k][0], kev[k][1])      # the walk continues: new frames, new poses
    barrier()
    fused_ms = statistics.mean(a.elapsed_time(b) for a, b in kev)
    return elapsed_ms, fused_ms, checksum


def pcie_ceiling(dev, barrier, h2d_bytes, d2h_bytes, reps=6):
    """What the host links allow for ONE e2e step of this rank, all ranks measuring together (whatever they share -- root
    complex, host memory -- is shared here as well), pinned cudaMemcpyAsync only:
      * `bound_ms`: the step's H2D bytes alone, nothing going the other way.  No pipeline can beat it: a true ceiling.
      * `mix_ms`:   the step's H2D and D2H bytes on two streams at once, until both are through: what a perfectly
                    overlapped pipeline with this byte mix would take if both copies started together (an estimate --
                    at N = 8 the real pipeline, whose D2H trails its H2D, has been measured slightly faster).
    Returns (bound_ms, mix_ms, H2D GB/s alone, H2D GB/s in the mix, D2H GB/s in the mix)."""
    import torch
    h2d_bytes, d2h_bytes = max(int(h2d_bytes), 1 << 20), max(int(d2h_bytes), 1 << 20)
    h_in = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(h2d_bytes, dtype=torch.uint8, device=dev)
    d_out = torch.empty(d2h_bytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def run(both):
        for _ in range(2):
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
            if both:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize(dev)
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(s1)
        b0.record(s2)
        for _ in range(reps):
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
            if both:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        a1.record(s1)
        b1.record(s2)
        torch.cuda.synchronize(dev)
        barrier()
        return a0.elapsed_time(a1), b0.elapsed_time(b1)

    ta_only, _ = run(False)
    ta, tb = run(True)
    return (ta_only / reps, max(ta, tb) / reps, reps * h2d_bytes / (ta_only * 1e6), reps * h2d_bytes / (ta * 1e6),
            reps * d2h_bytes / (tb * 1e6))


def run_e2e(args, n, dev, rank, barrier):
    """End to end through the reference-facing host-buffer entry (wsmg_map_update_host_ex): every step copies that
    step's NEW frames from pinned host memory, updates the maps and copies the ego maps back, inside the timed region."""
    import torch
    import wsmgmap_b200  # noqa: F401
    from wsmgmap_b200 import ops
    from wsmgmap_b200.synth import RandomWalk, make_depth
    s = SHAPE
    ne = min(n, args.e2e_envs)
    Ke = max(2, args.e2e_steps)
    de = ops.dims_for((ne, s["C"], s["Hf"], s["Wf"]), (ne, s["Hd"], s["Wd"], 1), ne, s["E"], s["G"], s["resolution"])
    pipe = ops.HostPipeline(de, dev, chunk_envs=args.e2e_chunk, zero_copy=args.e2e_mode == "zerocopy",
                            skip_dead_rows=args.e2e_mode == "rows")
    T = Ke + 1
    gen = torch.Generator().manual_seed(4321 + rank)
    feat_h = torch.empty(ne + T, s["C"], s["Hf"], s["Wf"]).pin_memory()          # sliding window: a new frame per env per step
    feat_h.uniform_(0, 1, generator=gen)
    cgen = torch.Generator().manual_seed(77 + rank)
    kinds = [make_depth(k, 8, s["Hd"], s["Wd"], cgen) for k in DEPTH_KINDS]
    depth_h = torch.stack([kinds[b % 4][(b // 4) % 8] for b in range(ne + T)], 0).contiguous().pin_memory()
    walk = RandomWalk(ne, seed=70 + rank)
    poses = [walk.step() for _ in range(T)]
    gps_h, comp_h, mask_h = (torch.stack([p[i] for p in poses]).contiguous().pin_memory() for i in range(3))
    ego_h = torch.empty(ne, s["C"], s["E"], s["E"]).pin_memory()
    gmap_e = torch.zeros(ne, s["G"], s["G"], s["C"], device=dev)
    stream = torch.cuda.current_stream(dev)
    pipe.step(feat_h[0:ne], depth_h[0:ne], gps_h[0], comp_h[0], mask_h[0], gmap_e, ego_h)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in range(1, Ke + 1):
        pipe.step(feat_h[k:k + ne], depth_h[k:k + ne], gps_h[k], comp_h[k], mask_h[k], gmap_e, ego_h)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    dense = ne * 4 * (s["C"] * s["Hf"] * s["Wf"] + s["Hd"] * s["Wd"] + 4)
    h2d = dense
    if args.e2e_mode == "rows":      # bytes that actually cross the bus: live feature rows + depth + pose (mean over the steps)
        rows = 0
        for k in range(1, Ke + 1):
            lo, hi = ops.host_live_rows(depth_h[k:k + ne], de)
            rows += int((hi - lo + 1).clamp(min=0).sum())
        h2d = int(4 * (rows / Ke * s["Wf"] * s["C"] + ne * (s["Hd"] * s["Wd"] + 4)))
    d2h = ne * 4 * s["C"] * s["E"] * s["E"]
    return dict(ms=ms, envs=ne, steps=Ke, h2d=h2d, d2h=d2h, dense=dense)


def policy_forward_record(dev):
    """BASELINE.json configs[4]: the reference's full policy forward at batch 64 with its own mapping module (on this
    GPU) and with the drop-in swapped in: ms per step and the map update's share of it."""
    import torch
    from baseline import policy_harness as ph
    if not ph.available():
        return {"unavailable": "reference policy files not staged under baseline/_ref (run __graft_entry__.build() where /root/reference exists)"}
    bs, warm, steps = 64, 2, 5
    frames = ph.make_observations(bs, warm + steps, seed=11, device=dev)
    rec = {"batch": bs, "steps": steps, "weights": "random (no checkpoints offline)", "third_party": "stand-ins for gym / habitat / habitat_baselines (baseline/habitat_shims.py)"}
    outs = {}
    for which in ("reference", "dropin"):
        policy = ph.build_policy(which, bs, dev, seed=7)
        ph.rollout(policy, frames[:warm])
        res, total, share = ph.rollout(policy, frames[warm:], time_it=True)
        outs[which] = res[-1][2]
        rec[which] = {"ms_per_step": total / steps, "map_update_ms_per_step": share / steps, "map_update_share": share / total}
        del policy
        torch.cuda.empty_cache()
    rec["speedup_whole_forward"] = rec["reference"]["ms_per_step"] / rec["dropin"]["ms_per_step"]
    return rec


def run_native(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import wsmgmap_b200  # noqa: F401
    from wsmgmap_b200 import _lib, ops, shard
    from wsmgmap_b200.synth import make_depth

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl native needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_node, numa_note = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    s = SHAPE
    cfg = workload_config(args, world)
    n = len(shard.shard_range(cfg["envs_total"], world, rank)) if cfg["scaling"] == "strong" else cfg["envs_per_gpu"]
    K, W = args.steps, max(args.warmup, 3)
    T = W + 2 * K

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---- device-resident throughput (`value`) and the dominant kernel alone (roofline) --------------------
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if n < 64 else None
    res = Resident(n, T, dev, rank, lib, nhwc=args.feat_layout == "nhwc")
    with ClockSampler(local_rank) as clk:
        elapsed_ms, fused_ms, checksum = time_resident(res, K, W, barrier, flush_buf)

    # ---- per depth distribution (uniform / near / room), short runs on up to 256 envs ----------
    by_depth = {}
    if not args.no_by_depth:
        nd = min(n, 256)
        cgen = torch.Generator().manual_seed(5 + rank)
        gmap_d = torch.zeros(nd, s["G"], s["G"], s["C"], device=dev)
        dd_ = ops.dims_for((nd, s["C"], s["Hf"], s["Wf"]), (nd, s["Hd"], s["Wd"], 1), nd, s["E"], s["G"], s["resolution"])
        P = lambda x: ctypes.c_void_p(x.data_ptr())  # noqa: E731
        sp = ctypes.c_void_p(res.stream.cuda_stream)
        for kind in DEPTH_KINDS:
            dk = torch.cat([make_depth(kind, 8, s["Hd"], s["Wd"], cgen)] * (nd // 8 + 1), 0)[:nd].to(dev).contiguous()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(6)]
            for k in range(len(ev) + 2):
                if k >= 2:
                    ev[k - 2][0].record(res.stream)
                rc = lib.wsmg_map_update(P(res.feat_all[k:k + nd]), P(dk), P(res.gps[k]), P(res.compass[k]), P(res.masks[k]), P(gmap_d),
                                         P(res.ego), None, P(res.scratch), res.scratch.numel(), ctypes.byref(dd_), sp)
                _lib.check(rc, "wsmg_map_update")
                if k >= 2:
                    ev[k - 2][1].record(res.stream)
            torch.cuda.synchronize(dev)
            ms = statistics.median(a.elapsed_time(b) for a, b in ev)
            _, inv = ops.unproject_index(dk[:8], s["Hf"], s["Wf"], s["E"], s["G"], s["resolution"])
            by_depth[kind] = {"frames_per_s_per_gpu": nd / (ms / 1e3), "envs": nd,
                              "writing_pixel_frac": float(1.0 - inv.float().mean())}
        del gmap_d
    del res
    torch.cuda.empty_cache()

    # ---- weak-scaling variant of the same workload (N > 1 only; at N = 1 it is the run above) --------------
    weak_ms = None
    if world > 1 and cfg["scaling"] == "strong" and not args.no_weak:
        resw = Resident(cfg["envs_total"], T, dev, rank, lib)
        weak_ms, _, _ = time_resident(resw, K, W, barrier)
        del resw
        torch.cuda.empty_cache()

    # ---- the named small-batch configuration (BASELINE configs[1], 8 envs, L2 flushed), single GPU ---------
    small = None
    if world == 1 and args.workload == "sweep1024" and not args.no_small_batch:
        fb = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        rs = Resident(8, T, dev, rank, lib)
        sm_ms, sm_fused, _ = time_resident(rs, K, W, barrier, fb)
        B8 = algorithmic_bytes_per_frame() * 8
        peak, _ = measured_peak()
        small = {"workload": "env8 (BASELINE configs[1]): 8 envs on one GPU, L2 flushed between steps", "envs": 8,
                 "ms_per_step": sm_ms / K, "frames_per_s": 8 * K / (sm_ms / 1e3), "kernel_ms": sm_fused,
                 "whole_step_frac": B8 / (sm_ms / K / 1e3) / 1e9 / peak, "kernel_frac": B8 / (sm_fused / 1e3) / 1e9 / peak}
        del rs, fb
        torch.cuda.empty_cache()

    # ---- end to end from pinned host buffers (`e2e`) and the link's ceiling measured in the same run --------
    e2e = run_e2e(args, n, dev, rank, barrier)
    link_ms, mix_ms, h2d_alone_gbs, h2d_gbs, d2h_gbs = pcie_ceiling(dev, barrier, e2e["h2d"], e2e["d2h"])

    # ---- comparators on one GPU: the reference's ops on this GPU, index flips against them, the policy forward ----
    torch_cuda, flips, policy_rec = None, None, None
    if world == 1 and not args.no_cpu_baseline:
        import warnings
        warnings.filterwarnings("ignore", message=".*align_corners.*")
        nb = CPU_SAMPLE_ENVS
        feat, depth, walk = _comparator_inputs(nb, 0)
        feat, depth = feat.to(dev), depth.to(dev)
        step, kind = reference_stepper(nb, dev)
        poses = [tuple(x.to(dev) for x in walk.step()) for _ in range(13)]
        with torch.no_grad():
            for k in range(3):
                step(feat, depth, *poses[k])
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            for k in range(10):
                step(feat, depth, *poses[3 + k])
            torch.cuda.synchronize(dev)
        torch_cuda = {"value": nb * 10 / (time.perf_counter() - t0), "unit": "frames/s", "kind": kind,
                      "sample": f"{nb} envs/step x 10 steps, the reference's PyTorch ops on the same B200 "
                                f"({'unmodified rgb_mapping.py' if kind == 'reference' else 'oracle port'})"}
        try:
            from oracle.cuda_flips import flips_against_cuda_reference
            fl = flips_against_cuda_reference()
            flips = {k: {"flipped_pixels": v[0], "pixels": v[1], "rate": v[0] / v[1], "max_abs_ego_diff": v[2], "max_abs_ego": v[3]}
                     for k, v in fl.items()}
            flips["note"] = ("cell-index changes of the kernels against the reference's torch ops run ON CUDA (8 envs per depth kind, "
                             "real shapes); against the reference on the CPU -- the parity target -- there are none")
        except Exception as e:                                   # a comparator must never cost the bench line
            flips = {"unavailable": f"{type(e).__name__}: {e}"}
        if not args.no_policy:
            try:
                policy_rec = policy_forward_record(dev)
            except Exception as e:
                policy_rec = {"unavailable": f"{type(e).__name__}: {e}"}

    # ---- gather (max over ranks) ---------------------------------------------------------------------------
    stats = torch.tensor([elapsed_ms, fused_ms, e2e["ms"], checksum, float(n), weak_ms or 0.0, h2d_gbs, d2h_gbs,
                          float(e2e["h2d"]), float(e2e["d2h"]), float(e2e["dense"]), link_ms, mix_ms, h2d_alone_gbs], dtype=torch.float64, device=dev)
    allst = shard.gather_stats(stats)             # the only collective: a few dozen bytes over NVLink
    if rank == 0:
        max_ms = float(allst[:, 0].max())
        max_fused = float(allst[:, 1].max())
        max_e2e = float(allst[:, 2].max())
        frames_per_step = float(allst[:, 4].sum())
        fps = shard.job_throughput(allst[:, 4] * K, allst[:, 0])
        B = algorithmic_bytes_per_frame()
        peak, peak_src = measured_peak()
        n_max = float(allst[:, 4].max())
        achieved = B * n_max / (max_fused / 1e3) / 1e9
        tr_env, tr_src = ncu_traffic_per_env()
        traffic = args.traffic if args.traffic is not None else (tr_env * n_max if tr_env is not None else None)
        ne, Ke = e2e["envs"], e2e["steps"]
        e2e_fps = ne * world * Ke / (max_e2e / 1e3)
        h2d_step, d2h_step = float(allst[:, 8].sum()), float(allst[:, 9].sum())
        link_h2d, link_d2h = float(allst[:, 6].min()), float(allst[:, 7].min())
        ceiling = ne * world / (float(allst[:, 11].max()) / 1e3)      # frames/s if a step were nothing but its H2D copies
        mix_fps = ne * world / (float(allst[:, 12].max()) / 1e3)      # ... its H2D and D2H copies, started together
        line = {
            "metric": "map-update frames/sec", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": max_ms / K, "higher_is_better": True, "scaling": cfg["scaling"],
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "native", "config": cfg,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": "given on the command line" if args.traffic is not None else
                         (f"committed ncu capture ({tr_src}) scaled to this launch's env count, not measured in this run" if tr_env is not None else None),
                         "kernel": "k_fused", "kernel_ms": max_fused,
                         "algorithmic_bytes_per_launch": B * n_max, "peak_source": peak_src,
                         "whole_step_frac": B * frames_per_step / (max_ms / K / 1e3) / 1e9 / (peak * world)},
            "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": h2d_step,
                    "d2h_bytes_per_step": d2h_step, "envs_per_gpu": ne, "steps": Ke,
                    "api": "wsmg_map_update_host_ex (pinned host buffers, chunked H2D/compute/D2H), a new frame per env per step",
                    "mode": args.e2e_mode, "numa_node_rank0": numa_node, "numa_note": numa_note,
                    "h2d_bytes_per_step_dense": float(allst[:, 10].sum()),
                    "pcie_ceiling_gbs": {"h2d_alone": float(allst[:, 13].min()), "h2d_in_mix": link_h2d, "d2h_in_mix": link_d2h,
                                         "how": "pinned cudaMemcpyAsync of one e2e step's bytes, all ranks together, GB/s = min over ranks: "
                                                "h2d_alone = the step's H2D bytes with nothing going back (the ceiling: no pipeline can beat it); "
                                                "in_mix = H2D and D2H bytes on two streams at once (estimate for a perfectly overlapped pipeline)"},
                    "pcie_ceiling_frames_per_s": ceiling, "frac_of_pcie_ceiling": e2e_fps / ceiling,
                    "pcie_mix_frames_per_s": mix_fps},
            "gpu_launches": 2 * K * world,               # k_cells (+ reset / rotation-setup blocks) and k_fused per step and rank
            "clocks": clk.summary(),
            "checksums": [float(x) for x in allst[:, 3]],
            "by_depth": by_depth,
        }
        if weak_ms is not None:
            wmax = float(allst[:, 5].max())
            line["weak"] = {"value": cfg["envs_total"] * world * K / (wmax / 1e3), "unit": "frames/s", "ms_per_step": wmax / K,
                            "envs_per_gpu": cfg["envs_total"], "scaling": "weak",
                            "note": "same run, every GPU owning the full 1024 envs (round 1's default)"}
        if small is not None:
            line["small_batch"] = small
        if torch_cuda is not None:
            line["torch_cuda_baseline"] = torch_cuda
        if flips is not None:
            line["cuda_reference_flips"] = flips
        if policy_rec is not None:
            line["policy_forward"] = policy_rec
        if world == 1 and not args.no_cpu_baseline:
            fps_cpu, cores, _, kind = cpu_reference_fps(CPU_SAMPLE_ENVS, args.cpu_steps, 2)
            what = "the unmodified reference file rgb_mapping.py" if kind == "reference" else "oracle port of the reference's torch CPU path"
            line["cpu_baseline"] = {"value": fps_cpu, "unit": "frames/s", "cores": cores, "kind": kind,
                                    "sample": f"{CPU_SAMPLE_ENVS} envs/step x {args.cpu_steps} steps (2 warm-up) of the same shapes and depth mix, {what}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="sweep1024", choices=["sweep1024", "env8"])
    ap.add_argument("--envs", type=int, default=0, help="override the env count (total when strong, per GPU when weak)")
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"],
                    help="strong (default for sweep1024): the env count in total, sharded; weak: that many per GPU")
    ap.add_argument("--e2e-envs", type=int, default=128, help="envs per GPU in the host-buffer (e2e) leg")
    ap.add_argument("--e2e-chunk", type=int, default=4, help="largest chunk of the host-buffer entry (it ramps up to and down from this size)")
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--e2e-mode", default="rows", choices=["copy", "rows", "zerocopy"],
                    help="host-buffer leg: stage the whole feature tensor, only the rows that hold a writing pixel, "
                         "or let the scatter read the pinned buffer")
    ap.add_argument("--cpu-steps", type=int, default=120, help="steps of the CPU baseline sample (8 envs each, ~10-20 s)")
    ap.add_argument("--traffic", type=float, default=None, help="dram bytes per k_fused launch from ncu, if known")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip every comparator (CPU, torch-on-CUDA, flips, policy)")
    ap.add_argument("--no-by-depth", action="store_true", help="skip the per-depth-distribution runs")
    ap.add_argument("--no-weak", action="store_true", help="skip the weak-scaling variant at N > 1")
    ap.add_argument("--no-small-batch", action="store_true", help="skip the env8 sub-record")
    ap.add_argument("--no-policy", action="store_true", help="skip the policy-forward record")
    ap.add_argument("--feat-layout", default="nchw", choices=["nchw", "nhwc"],
                    help="memory layout of the resident feature tensor (nhwc: a channels_last producer, consumed without a permute copy)")
    ap.add_argument("--shape", default="real", choices=sorted(SHAPES), help="tensor shapes (secondary shapes of SURVEY 8d)")
    args = ap.parse_args()
    SHAPE.clear()
    SHAPE.update(SHAPES[args.shape])
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1:
        # not under torchrun: relaunch ourselves one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_native(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
