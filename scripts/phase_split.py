"""Time the stage entry points against the whole step at 1024 envs (device-resident):
scatter-only (k_cells + k_fused scatter + proj dump) and registration-only (proj load + phases 2-4)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
# the phase-skipping switch exists only in the profiling build: python ws-mgmap_b200/build.py --phase-skip
os.environ.setdefault("WSMG_LIB_PATH", os.path.join(ROOT, "ws-mgmap_b200", "lib", "libwsmg_phaseskip.so"))
import torch
import wsmgmap_b200
from wsmgmap_b200 import ops
from wsmgmap_b200.synth import make_depth

dev = torch.device("cuda", 0)
n, c = 1024, 64
gen = torch.Generator(device=dev).manual_seed(0)
feat = torch.rand(n, c, 224, 224, generator=gen, device=dev)
cg = torch.Generator().manual_seed(1)
kinds = [make_depth(k, 8, 256, 256, cg) for k in ("uniform", "near", "room2", "room4")]
depth = torch.stack([kinds[b % 4][(b // 4) % 8] for b in range(n)], 0).to(dev).contiguous()
gps = torch.randn(n, 2, device=dev); compass = torch.rand(n, 1, device=dev) * 6 - 3
ones = torch.ones(n, 1, device=dev)
gmap = torch.zeros(n, 240, 240, c, device=dev)
d = ops.dims_for(feat.shape, depth.shape, n)
scratch = ops.alloc_scratch(d, dev)
ego = torch.empty(n, c, 100, 100, device=dev)

def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

t_full = timeit(lambda: ops.map_update(feat, depth, gps, compass, ones, gmap, scratch=scratch, ego=ego))
proj = ops.scatter_max(feat, depth)
t_scat = timeit(lambda: ops.scatter_max(feat, depth))
t_reg = timeit(lambda: ops.register_fuse_retrieve(proj, gps, compass, ones, gmap))
dead = torch.ones_like(depth)            # 10 m everywhere: no pixel writes, no feature is read
t_dead = timeit(lambda: ops.map_update(feat, dead, gps, compass, ones, gmap, scratch=scratch, ego=ego))
for mask, name in ((1, "scatter"), (2, "first rotation"), (4, "band loop (incl. its TMA traffic)"), (8, "output rotation"), (15, "all four"),
                   (14, "all but the scatter"), (7, "all but the output rotation"), (16, "band loop: crop (3b) only"),
                   (32, "band loop: fuse (3a) only"), (48, "band loop: crop + fuse, keep TMA + barriers"),
                   (64, "band loop: the TMA-arrival wait only"), (112, "band loop: crop + fuse + wait (barriers + TMA loads remain)"),
                   (15 + 512, "all four + key decode"), (15 + 1024, "all four + translation tables"),
                   (15 + 2048, "all four + key-plane init"), (15 + 512 + 1024 + 2048, "all four + decode + tables + init"),
                   (16384, "scatter: the shared-memory atomics only"), (32768, "scatter: the feature copies only (cp.async)"),
                   (16384 + 32768, "scatter: atomics + copies (codes, staging reads, run merging remain)"),
                   (65536, "output rotation: the B tile copies only (cp.async from the crop slot)"),
                   (8 + 65536, "output rotation incl. its tile copies"),
                   (4096, "everything: items return at entry (k_reset + k_cells + launch + dispenser cost)")):
    os.environ["WSMG_DEBUG_SKIP"] = str(mask)
    t = timeit(lambda: ops.map_update(feat, depth, gps, compass, ones, gmap, scratch=scratch, ego=ego))
    print(f"  skip {name:34s} -> {t:.3f} ms  (saves {t_full - t:+.3f})")
os.environ.pop("WSMG_DEBUG_SKIP")
print(f"whole step {t_full:.3f} ms | scatter-only (+2.56 MB/env proj write) {t_scat:.3f} ms | registration-only (+proj read) {t_reg:.3f} ms | "
      f"whole step with an all-dead depth (no feature reads, empty fan) {t_dead:.3f} ms")
