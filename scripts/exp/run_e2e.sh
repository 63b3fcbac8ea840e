cd $GRAFT_REPO_ROOT
for c in 2 4 8 16 32; do echo "== e2e chunk $c"; python bench.py --envs 128 --steps 3 --warmup 3 --e2e-chunk $c --no-cpu-baseline --no-by-depth --no-small-batch 2>/dev/null | python -c '
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); e = d["e2e"]
print("e2e", round(e["value"]), "ceiling", round(e["pcie_ceiling_frames_per_s"]), "frac", round(e["frac_of_pcie_ceiling"], 3), "h2d GB/s", round(e["pcie_ceiling_gbs"]["h2d"], 1))'; done
