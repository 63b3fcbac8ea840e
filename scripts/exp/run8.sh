#!/bin/bash
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest gpu"; timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; grep -E "^(E   |FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | cut -c1-300 | tail -8
for envs in 1024 128; do
echo "== bench envs=$envs"; timeout 900 python bench.py --envs $envs --no-cpu-baseline --no-by-depth --no-small-batch --steps 30 2>gpurun_out/bench.err | tee gpurun_out/bench_$envs.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',round(d['value']),'ms/step',round(d['ms_per_step'],4),'k_fused ms',round(d['roofline']['kernel_ms'],4),'frac',round(d['roofline']['frac'],3),'e2e',round(d['e2e']['value']))"; tail -2 gpurun_out/bench.err
done
echo "== bench env8"; timeout 600 python bench.py --workload env8 --no-cpu-baseline --no-by-depth 2>>gpurun_out/bench.err | tee gpurun_out/bench_env8.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('env8 us/step',round(1e3*d['ms_per_step'],1),'k_fused us',round(1e3*d['roofline']['kernel_ms'],1))"
echo "== host overhead"; timeout 300 python scripts/host_overhead.py 2>&1 | tail -3
echo "== racecheck"; timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -q -x -k golden_trajectory 2>&1 | tail -5 | tee gpurun_out/racecheck.log
