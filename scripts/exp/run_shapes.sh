cd $GRAFT_REPO_ROOT
fmt='import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d["config"]["workload"][:60], "| ms_per_step", round(d["ms_per_step"],4), "value", round(d["value"]), "kernel_ms", round(d["roofline"]["kernel_ms"],4), "frac", round(d["roofline"]["frac"],3), "whole", round(d["roofline"]["whole_step_frac"],3))'
for e in 128 256 512; do echo "== envs $e"; python bench.py --envs $e --no-cpu-baseline --no-by-depth --no-small-batch --e2e-envs 16 --e2e-steps 2 2>/dev/null | tee gpurun_out/bench_envs$e.json | python -c "$fmt"; done
for sh in b256 b256c27; do echo "== shape $sh"; python bench.py --shape $sh --envs 512 --no-cpu-baseline --no-by-depth --no-small-batch --e2e-envs 16 --e2e-steps 2 2>/dev/null | tee gpurun_out/bench_$sh.json | python -c "$fmt"; done
echo "== nhwc"; python bench.py --feat-layout nhwc --no-cpu-baseline --no-by-depth --no-small-batch --e2e-envs 16 --e2e-steps 2 2>/dev/null | tee gpurun_out/bench_nhwc.json | python -c "$fmt"
echo "== host overhead"; python scripts/host_overhead.py 2>&1 | tail -6
