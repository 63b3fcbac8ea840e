cd $GRAFT_REPO_ROOT
fmt='
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); e = d["e2e"]
print("e2e", round(e["value"]), "bound", round(e["pcie_ceiling_frames_per_s"]), "mix", round(e["pcie_mix_frames_per_s"]), "frac", round(e["frac_of_pcie_ceiling"], 3), "of mix", round(e["value"]/e["pcie_mix_frames_per_s"],3), "MB/frame", round(e["h2d_bytes_per_step"]/e["envs_per_gpu"]/1e6,2))'
for args in "--e2e-mode rows" "--e2e-mode copy" "--e2e-mode rows --feat-layout nhwc" "--e2e-mode copy --e2e-chunk 16" "--e2e-mode rows --e2e-chunk 2"; do echo "== $args"; python bench.py --envs 128 --steps 3 --warmup 3 $args --no-cpu-baseline --no-by-depth --no-small-batch 2>/dev/null | python -c "$fmt"; done
