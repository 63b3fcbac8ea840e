"""Host-side cost per map update at the reference's batch size (8 envs): the drop-in module call,
the ops wrapper, and the bare C ABI call (CPU time per call, GPU queue kept non-empty)."""
import ctypes, os, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wsmgmap_b200
from wsmgmap_b200 import _lib, ops
from wsmgmap_b200.rgb_mapping import RGBMapping

dev = torch.device("cuda", 0)
bs, c = 8, 64
feat = torch.rand(bs, c, 224, 224, device=dev); depth = torch.rand(bs, 256, 256, 1, device=dev) * 0.6
gps = torch.zeros(bs, 2, device=dev); compass = torch.zeros(bs, 1, device=dev); masks = torch.ones(bs, 1, device=dev)
m = RGBMapping(types.SimpleNamespace(gpu_id=0, num_proc=bs, resolution=0.12, egocentric_map_size=100, global_map_size=240, map_depth=c))
lib = _lib.load()
d = ops.dims_for(feat.shape, depth.shape, bs)
scratch = ops.alloc_scratch(d, dev); ego = torch.empty(bs, c, 100, 100, device=dev)
P = lambda t: ctypes.c_void_p(t.data_ptr())
sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
def t_module():
    m(feat, dict(depth=depth, gps=gps, compass=compass), masks)
def t_ops():
    ops.map_update(feat, depth, gps, compass, masks, m.full_global_map, scratch=scratch, ego=ego)
def t_cabi():
    lib.wsmg_map_update(P(feat), P(depth), P(gps), P(compass), P(masks), P(m.full_global_map), P(ego), None, P(scratch), scratch.numel(), ctypes.byref(d), sp)
for name, fn in (("module.forward", t_module), ("ops.map_update", t_ops), ("C ABI (ctypes)", t_cabi)):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    n = 300
    t0 = time.perf_counter()
    for _ in range(n): fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{name:18s} host {1e6*(t1-t0)/n:7.1f} us/call   wall incl. GPU drain {1e6*(t2-t0)/n:7.1f} us/call")
