#!/usr/bin/env python
"""Benchmark of the map update (BASELINE.json metric: map-update frames/sec at 1/2/4/8 B200 +
achieved HBM GB/s vs peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--workload sweep1024|env8]

Workload `sweep1024` (BASELINE configs[3], the configuration the metric is quoted on): 1024 envs at the
reference shapes (C=64 features 224x224, depth 256x256, ego 100, global 240), sharded 1024/N per GPU,
no cross-GPU traffic in the loop; one NCCL all_gather of the per-rank stats at the end.  A step is one
map update of every env (1 frame = 1 env x 1 step).  Inputs (13 GB of features per 1024 envs) are far
larger than the 126 MB L2, so no L2 flush is needed between iterations.
`env8` is BASELINE configs[1] (8 envs on one GPU; L2 flushed between steps).

One JSON line on stdout (rank 0).  --impl reference times the oracle port of the reference's PyTorch CPU
path on the host cores (the reference is Python + Habitat; only this file of it runs without a simulator).
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
    written by scripts/ncu_summary.sh); None when no capture is committed."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return float(json.load(open(p))["k_fused_dram_bytes_per_env"])
    except Exception:
        return None


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


def cpu_reference_fps(n_envs, steps, warmup, seed=0):
    """The reference's PyTorch CPU path (oracle port, asserted equal to the reference file by the tests),
    all host threads, on `n_envs` envs per step of the same synthetic workload."""
    import torch
    from oracle.mapping_oracle import OracleMapper
    import wsmgmap_b200  # noqa: F401
    from wsmgmap_b200.synth import RandomWalk, make_depth, make_features
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    s = SHAPE
    gen = torch.Generator().manual_seed(seed)
    feat = make_features(n_envs, s["C"], s["Hf"], s["Wf"], gen)
    depth = torch.cat([make_depth(DEPTH_KINDS[b % 4], 1, s["Hd"], s["Wd"], gen) for b in range(n_envs)], 0)
    walk = RandomWalk(n_envs, seed=seed)
    orc = OracleMapper(n_envs, s["C"])
    times = []
    for t in range(warmup + steps):
        gps, compass, masks = walk.step()
        t0 = time.perf_counter()
        orc.step(feat, depth, gps, compass, masks)
        dt = time.perf_counter() - t0
        if t >= warmup:
            times.append(dt)
    total = sum(times)
    return n_envs * len(times) / total, cores, total / len(times)


def run_reference(args, rank, world):
    if rank != 0:
        return
    n = 8
    fps, cores, sec = cpu_reference_fps(n, args.steps, args.warmup)
    sample = f"{n} envs/step x {args.steps} steps of the {args.workload} workload (mixed depth kinds), torch CPU"
    line = {
        "metric": "map-update frames/sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(local_rank):
    """Run this rank on the CPUs of the NUMA node its GPU hangs off, so that the pinned host buffers of the e2e leg are
    first-touched in that node's memory (what a production launcher does with numactl).  Best effort; returns the node."""
    try:
        import torch
        bdf = torch.cuda.get_device_properties(local_rank).pci_bus_id if hasattr(
            torch.cuda.get_device_properties(local_rank), "pci_bus_id") else None
        if bdf is None:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
            bdf = pynvml.nvmlDeviceGetPciInfo(h).busId
            bdf = bdf.decode() if isinstance(bdf, bytes) else bdf
        bdf = bdf.lower()
        if len(bdf.split(":")[0]) == 8:
            bdf = bdf[4:]
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.extend(range(int(a), int(b or a) + 1))
        allowed = set(os.sched_getaffinity(0)) & set(cpus)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None


def workload_config(args, world):
    """sweep1024 / env8: 1024 / 8 envs PER GPU by default (weak scaling -- every rank owns its envs and their map state,
    exactly like a rank of the reference owns its NUM_PROCESSES envs); --scaling strong keeps the total fixed and
    shards it, the literal reading of BASELINE.json configs[3]."""
    base = 1024 if args.workload == "sweep1024" else 8
    if args.envs:
        base = args.envs
    if args.scaling == "weak":
        per_gpu, total = base, base * world
    else:
        per_gpu, total = base // world, base
    return {"workload": f"{args.workload}: {per_gpu} envs per GPU, {total} in total ({args.scaling} scaling), env-sharded, "
                        f"C={SHAPE['C']} feat {SHAPE['Hf']}x{SHAPE['Wf']} depth {SHAPE['Hd']}x{SHAPE['Wd']} ego 100 global 240 fp32, depth kinds mixed "
                        f"{'/'.join(DEPTH_KINDS)}, random-walk poses, a new frame per env per step, masks=1 after the first step",
            "envs_total": total, "envs_per_gpu": per_gpu,
            "l2": "inputs larger than L2 (no flush)" if per_gpu >= 64 else "L2 flushed between steps",
            "bytes_per_frame_algorithmic": algorithmic_bytes_per_frame()}


def run_native(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import wsmgmap_b200  # noqa: F401
    from wsmgmap_b200 import _lib, ops
    from wsmgmap_b200.synth import RandomWalk, make_depth

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl native needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_node = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    s = SHAPE
    cfg = workload_config(args, world)
    n = cfg["envs_per_gpu"]
    K, W = args.steps, max(args.warmup, 3)
    flush_l2 = n < 64

    # ---- synthetic inputs, resident in HBM --------------------------------------------------
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    # Every env sees a NEW frame every step, as in a rollout: n + T frames are resident and step t reads the
    # window [t, t+n).  (Re-feeding the same frame would leave the max-fused map unchanged after a few steps,
    # and a map update that changes nothing also writes nothing.)
    T = W + 2 * K
    feat_all = torch.rand(n + T, s["C"], s["Hf"], s["Wf"], generator=gen, device=dev)
    cgen = torch.Generator().manual_seed(99 + rank)
    kinds = [make_depth(k, 8, s["Hd"], s["Wd"], cgen) for k in DEPTH_KINDS]
    depth_all = torch.stack([kinds[b % 4][(b // 4) % 8] for b in range(n + T)], 0).to(dev).contiguous()
    feat, depth = feat_all[:n], depth_all[:n]
    walk = RandomWalk(n, seed=7 + rank)
    poses = [walk.step() for _ in range(T)]
    gps = torch.stack([p[0] for p in poses]).to(dev)
    compass = torch.stack([p[1] for p in poses]).to(dev)
    masks = torch.stack([p[2] for p in poses]).to(dev)
    gmap = torch.zeros(n, s["G"], s["G"], s["C"], device=dev)
    ego = torch.empty(n, s["C"], s["E"], s["E"], device=dev)
    d = ops.dims_for(feat.shape, depth.shape, n, s["E"], s["G"], s["resolution"])
    scratch = ops.alloc_scratch(d, dev)
    stream = torch.cuda.current_stream(dev)
    sp = ctypes.c_void_p(stream.cuda_stream)
    P = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if flush_l2 else None

    def step(t, ev0=None, ev1=None):
        feat, depth = feat_all[t:t + n], depth_all[t:t + n]
        if ev0 is None:
            rc = lib.wsmg_map_update(P(feat), P(depth), P(gps[t]), P(compass[t]), P(masks[t]), P(gmap), P(ego), None,
                                     P(scratch), scratch.numel(), ctypes.byref(d), sp)
        else:
            rc = lib.wsmg_map_update_timed(P(feat), P(depth), P(gps[t]), P(compass[t]), P(masks[t]), P(gmap), P(ego),
                                           None, P(scratch), scratch.numel(), ctypes.byref(d), sp,
                                           ctypes.c_void_p(ev0.cuda_event), ctypes.c_void_p(ev1.cuda_event))
        _lib.check(rc, "wsmg_map_update")

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---- device-resident throughput (`value`) -----------------------------------------------
    for t in range(W):
        step(t)
    barrier()
    with ClockSampler(local_rank) as clk:
        if flush_l2:
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
            for k in range(K):
                flush_buf.fill_(k & 0xFF)
                evs[k][0].record(stream)
                step(W + k)
                evs[k][1].record(stream)
            barrier()
            elapsed_ms = sum(a.elapsed_time(b) for a, b in evs)
        else:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for k in range(K):
                step(W + k)
            e1.record(stream)
            barrier()
            elapsed_ms = e0.elapsed_time(e1)
    checksum = float(ego.double().sum().item()) + float(gmap[0].double().sum().item())

    # ---- the dominant kernel alone (roofline): events around k_fused inside the same step loop --
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for a, b in kev:
        a.record(stream)
        b.record(stream)       # materialise the handles
    torch.cuda.synchronize(dev)
    for k in range(K):
        if flush_l2:
            flush_buf.fill_(k & 0xFF)
        step(W + K + k, kev[k][0], kev[k][1])      # the walk continues: new frames, new poses
    barrier()
    fused_ms = statistics.mean(a.elapsed_time(b) for a, b in kev)

    # ---- end to end from pinned host buffers (`e2e`) ----------------------------------------
    ne = min(n, args.e2e_envs)
    de = ops.dims_for((ne,) + tuple(feat.shape[1:]), depth[:ne].shape, ne, s["E"], s["G"], s["resolution"])
    pipe = ops.HostPipeline(de, dev, chunk_envs=args.e2e_chunk, zero_copy=args.e2e_mode == "zerocopy",
                            skip_dead_rows=args.e2e_mode == "rows")
    feat_h = feat[:ne].cpu().pin_memory()
    depth_h = depth[:ne].cpu().pin_memory()
    gps_h, comp_h, mask_h = (x[:, :ne].contiguous().cpu().pin_memory() for x in (gps, compass, masks))
    ego_h = torch.empty(ne, s["C"], s["E"], s["E"]).pin_memory()
    gmap_e = torch.zeros(ne, s["G"], s["G"], s["C"], device=dev)
    Ke = max(2, min(K, args.e2e_steps))
    pipe.step(feat_h, depth_h, gps_h[0], comp_h[0], mask_h[0], gmap_e, ego_h)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in range(Ke):
        pipe.step(feat_h, depth_h, gps_h[1 + k], comp_h[1 + k], mask_h[1 + k], gmap_e, ego_h)
    e1.record(stream)
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    h2d = ne * 4 * (s["C"] * s["Hf"] * s["Wf"] + s["Hd"] * s["Wd"] + 4)
    if args.e2e_mode == "rows":      # bytes that actually cross the bus: live feature rows + depth + pose
        lo, hi = ops.host_live_rows(depth_h, de)
        rows = int((hi - lo + 1).clamp(min=0).sum())
        h2d = 4 * (rows * s["Wf"] * s["C"] + ne * (s["Hd"] * s["Wd"] + 4))
    d2h = ne * 4 * s["C"] * s["E"] * s["E"]

    # ---- per depth distribution (uniform / near / room), short runs on up to 256 envs ----
    by_depth = {}
    if not args.no_by_depth:
        nd = min(n, 256)
        gmap_d = torch.zeros(nd, s["G"], s["G"], s["C"], device=dev)
        dd_ = ops.dims_for((nd,) + tuple(feat.shape[1:]), depth[:nd].shape, nd, s["E"], s["G"], s["resolution"])
        for kind in DEPTH_KINDS:
            dk = torch.cat([make_depth(kind, 8, s["Hd"], s["Wd"], cgen)] * (nd // 8 + 1), 0)[:nd].to(dev).contiguous()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(6)]
            for k in range(len(ev) + 2):
                if k >= 2:
                    ev[k - 2][0].record(stream)
                rc = lib.wsmg_map_update(P(feat_all[k:k + nd]), P(dk), P(gps[k]), P(compass[k]), P(masks[k]), P(gmap_d), P(ego), None,
                                         P(scratch), scratch.numel(), ctypes.byref(dd_), sp)
                _lib.check(rc, "wsmg_map_update")
                if k >= 2:
                    ev[k - 2][1].record(stream)
            torch.cuda.synchronize(dev)
            ms = statistics.median(a.elapsed_time(b) for a, b in ev)
            _, inv = ops.unproject_index(dk[:8], s["Hf"], s["Wf"], s["E"], s["G"], s["resolution"])
            by_depth[kind] = {"frames_per_s_per_gpu": nd / (ms / 1e3), "envs": nd,
                              "writing_pixel_frac": float(1.0 - inv.float().mean())}
        del gmap_d

    # ---- the reference's PyTorch ops on this GPU (oracle port, device=cuda): the stock comparator ----
    torch_cuda_fps = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle.mapping_oracle import OracleMapper
        nb = 8
        orc = OracleMapper(nb, s["C"], device=dev)
        gp, cp_, mk = gps[:, :nb].contiguous(), compass[:, :nb].contiguous(), masks[:, :nb].contiguous()
        for k in range(3):
            orc.step(feat[:nb], depth[:nb], gp[k], cp_[k], mk[k])
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for k in range(10):
            orc.step(feat[:nb], depth[:nb], gp[3 + k], cp_[3 + k], mk[3 + k])
        torch.cuda.synchronize(dev)
        torch_cuda_fps = nb * 10 / (time.perf_counter() - t0)

    # ---- gather (max over ranks) ------------------------------------------------------------
    stats = torch.tensor([elapsed_ms, fused_ms, e2e_ms, checksum], dtype=torch.float64, device=dev)
    if world > 1:
        allst = [torch.zeros_like(stats) for _ in range(world)]
        dist.all_gather(allst, stats)             # the only collective: a few dozen bytes over NVLink
        allst = torch.stack(allst).cpu()
    else:
        allst = stats.cpu().unsqueeze(0)
    if rank == 0:
        max_ms = float(allst[:, 0].max())
        max_fused = float(allst[:, 1].max())
        max_e2e = float(allst[:, 2].max())
        frames = n * world * K
        fps = frames / (max_ms / 1e3)
        B = algorithmic_bytes_per_frame()
        peak, peak_src = measured_peak()
        achieved = B * n / (max_fused / 1e3) / 1e9
        tr_env = ncu_traffic_per_env()
        traffic = args.traffic if args.traffic is not None else (tr_env * n if tr_env is not None else None)
        line = {
            "metric": "map-update frames/sec", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": max_ms / K, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "native", "config": cfg,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "k_fused", "kernel_ms": max_fused,
                         "algorithmic_bytes_per_launch": B * n, "peak_source": peak_src,
                         "whole_step_frac": B * n * world / (max_ms / K / 1e3) / 1e9 / (peak * world)},
            "e2e": {"value": ne * world * Ke / (max_e2e / 1e3), "unit": "frames/s", "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": d2h * world, "envs_per_gpu": ne, "steps": Ke,
                    "api": "wsmg_map_update_host_ex (pinned host buffers, chunked H2D/compute/D2H)", "mode": args.e2e_mode,
                    "numa_node_rank0": numa_node,
                    "h2d_bytes_per_step_dense": ne * 4 * (s["C"] * s["Hf"] * s["Wf"] + s["Hd"] * s["Wd"] + 4) * world},
            "gpu_launches": 3 * K * world,
            "clocks": clk.summary(),
            "checksums": [float(x) for x in allst[:, 3]],
            "by_depth": by_depth,
        }
        if torch_cuda_fps is not None:
            line["torch_cuda_baseline"] = {"value": torch_cuda_fps, "unit": "frames/s", "kind": "port",
                                           "sample": "8 envs/step x 10 steps, the reference's PyTorch ops (oracle port) on the same B200"}
        if world == 1 and not args.no_cpu_baseline:
            fps_cpu, cores, _ = cpu_reference_fps(8, 24, 2)
            line["cpu_baseline"] = {"value": fps_cpu, "unit": "frames/s", "cores": cores, "kind": "port",
                                    "sample": "8 envs/step x 24 steps (2 warm-up) of the same shapes, oracle port of the reference's torch CPU path"}
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
    ap.add_argument("--envs", type=int, default=0, help="override the env count (per GPU when weak, total when strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: the workload's envs per GPU (default); strong: that many in total, sharded")
    ap.add_argument("--e2e-envs", type=int, default=128, help="envs per GPU in the host-buffer (e2e) leg")
    ap.add_argument("--e2e-chunk", type=int, default=16)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--e2e-mode", default="rows", choices=["copy", "rows", "zerocopy"],
                    help="host-buffer leg: stage the whole feature tensor, only the rows that hold a writing pixel, "
                         "or let the scatter read the pinned buffer")
    ap.add_argument("--traffic", type=float, default=None, help="dram bytes per k_fused launch from ncu, if known")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-by-depth", action="store_true", help="skip the per-depth-distribution runs")
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
