import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import wsmgmap_b200
from wsmgmap_b200 import ops
g = np.load("tests/golden/traj_small.npz")
DEV = "cuda:0"
feat = torch.from_numpy(g["feat0"]).to(DEV); depth = torch.from_numpy(g["depth0"]).unsqueeze(-1).contiguous().to(DEV)
print("unproject"); lin, inv = ops.unproject_index(depth, 56, 56); torch.cuda.synchronize(); print("ok")
print("scatter"); proj = ops.scatter_max(feat, depth); torch.cuda.synchronize(); print("ok", np.array_equal(proj.cpu().numpy(), g["proj0"]))
gmap = torch.zeros(2, 240, 240, 4, device=DEV)
print("update"); ego = ops.map_update(feat, depth, torch.from_numpy(g["gps0"]).to(DEV), torch.from_numpy(g["compass0"]).to(DEV), torch.from_numpy(g["masks0"]).to(DEV), gmap); torch.cuda.synchronize(); print("ok")
