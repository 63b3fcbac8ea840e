#!/bin/bash
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; grep -E "^(E   |FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | cut -c1-300 | tail -12
echo "== bench"; timeout 900 python bench.py --no-cpu-baseline --steps 20 2>gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-1500; tail -3 gpurun_out/bench.err
echo "== bench env8"; timeout 600 python bench.py --workload env8 --no-cpu-baseline --no-by-depth 2>>gpurun_out/bench.err | tee gpurun_out/bench_env8.json | cut -c1-900
