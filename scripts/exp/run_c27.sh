cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_parity.py -x -q -k "other_geometries or kernel_variants" 2>&1 | tail -2
python bench.py --shape b256c27 --envs 512 --no-cpu-baseline --no-by-depth --no-small-batch --e2e-envs 16 --e2e-steps 2 2>/dev/null | tee gpurun_out/bench_b256c27.json | python -c '
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print("ms_per_step", round(d["ms_per_step"],4), "value", round(d["value"]), "kernel_ms", round(d["roofline"]["kernel_ms"],4), "frac", round(d["roofline"]["frac"],3), "whole", round(d["roofline"]["whole_step_frac"],3))'
