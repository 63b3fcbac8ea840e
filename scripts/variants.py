"""Kernel-variant A/B timing in ONE gpurun call.
Build here (no GPU needed):   python scripts/variants.py build name:DEF1,DEF2 name2:DEF ...
Run on the GPU box:           python scripts/variants.py run [skipmask ...]
Every variant is a profiling build (-DWSMG_PHASE_SKIP) so the same phase-skip masks apply to all of them; each
(variant, mask) pair is timed in its own process (the library is loaded once per process)."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIBDIR = os.path.join(ROOT, "ws-mgmap_b200", "lib")

def build(specs):
    import wsmgmap_b200  # noqa: F401
    from wsmgmap_b200.build import build_cuda
    for f in glob.glob(os.path.join(LIBDIR, "libwsmg_var_*.so*")):
        os.remove(f)
    for spec in specs:
        name, _, rest = spec.partition(":")          # name:DEF1,DEF2[;nvcc-flag;nvcc-flag]
        defs, _, fl = rest.partition(";")
        print(build_cuda(force=True, phase_skip=True, variant="var_" + name, defines=[d for d in defs.split(",") if d],
                         flags=[f for f in fl.split(";") if f]))

def child(envs):
    import torch
    import wsmgmap_b200  # noqa: F401
    from wsmgmap_b200 import ops
    from wsmgmap_b200.synth import make_depth
    dev = torch.device("cuda", 0)
    n, c = envs, 64
    gen = torch.Generator(device=dev).manual_seed(0)
    feat = torch.rand(n, c, 224, 224, generator=gen, device=dev)
    cg = torch.Generator().manual_seed(1)
    kinds = [make_depth(k, 8, 256, 256, cg) for k in ("uniform", "near", "room2", "room4")]
    depth = torch.stack([kinds[b % 4][(b // 4) % 8] for b in range(n)], 0).to(dev).contiguous()
    gps = torch.randn(n, 2, device=dev, generator=gen); compass = torch.rand(n, 1, device=dev, generator=gen) * 6 - 3
    ones = torch.ones(n, 1, device=dev)
    gmap = torch.zeros(n, 240, 240, c, device=dev)
    d = ops.dims_for(feat.shape, depth.shape, n)
    scratch = ops.alloc_scratch(d, dev)
    ego = torch.empty(n, c, 100, 100, device=dev)
    fn = lambda: ops.map_update(feat, depth, gps, compass, ones, gmap, scratch=scratch, ego=ego)
    if os.environ.get("VAR_STAGE") == "scatter":
        fn = lambda: ops.scatter_max(feat, depth)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 10)
    print(f"{best:.3f}", float(ego.double().sum()) if os.environ.get("WSMG_DEBUG_SKIP", "0") == "0" else "")

if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    elif sys.argv[1] == "child":
        child(int(sys.argv[2]))
    else:
        masks = sys.argv[2:] or ["0"]
        envs = os.environ.get("VAR_ENVS", "1024")
        libs = sorted(glob.glob(os.path.join(LIBDIR, "libwsmg_var_*.so")))
        print("variant".ljust(28) + "".join(f"skip={m}".rjust(14) for m in masks))
        for lib in libs:
            row = os.path.basename(lib)[len("libwsmg_var_"):-3].ljust(28)
            for m in masks:
                env = dict(os.environ, WSMG_LIB_PATH=lib, WSMG_DEBUG_SKIP=m)
                r = subprocess.run([sys.executable, __file__, "child", envs], env=env, capture_output=True, text=True)
                out = r.stdout.strip().split()
                row += (out[0] if out else "ERR").rjust(14)
                if m == "0" and len(out) > 1: row += f" [{out[1][:14]}]"
            print(row, flush=True)
