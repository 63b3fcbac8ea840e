#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; grep -E "^(E   |FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | cut -c1-300 | tail -8
for lay in nchw nhwc; do
echo "== bench $lay"; timeout 900 python bench.py --feat-layout $lay --no-cpu-baseline --no-by-depth --no-small-batch --steps 20 2>gpurun_out/bench.err | tee gpurun_out/bench_$lay.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',round(d['value']),'ms/step',round(d['ms_per_step'],4),'k_fused ms',round(d['roofline']['kernel_ms'],4),'frac',round(d['roofline']['frac'],3),'e2e',round(d['e2e']['value']))"; tail -2 gpurun_out/bench.err
done
