#!/bin/bash
# Short gpurun call for kernel iteration: GPU parity tests of the kernels, the device-resident bench (1024 envs and 8 envs),
# and the phase split of the profiling build when it was shipped.  Usage: bash scripts/gpu_quick.sh [tests|notests]
OUT=gpurun_out
mkdir -p $OUT
if [ "${1:-tests}" = "tests" ]; then
  echo "== pytest gpu (parity)"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -x -q > $OUT/pytest_gpu.log 2>&1
  grep -E "^(E   |FAILED|ERROR)|passed|failed" $OUT/pytest_gpu.log | cut -c1-300 | tail -12
fi
echo "== bench"; timeout 600 python bench.py --no-cpu-baseline --no-by-depth --e2e-envs 16 --e2e-steps 2 2>$OUT/bench.err | tee $OUT/bench_quick.json | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms_per_step', d['ms_per_step'], 'value', d['value'], 'kernel_ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], 'small', d.get('small_batch', {}).get('ms_per_step'), d.get('small_batch', {}).get('kernel_ms'))"
tail -3 $OUT/bench.err
[ -f ws-mgmap_b200/lib/libwsmg_phaseskip.so ] && { echo "== phase split"; timeout 600 python scripts/phase_split.py > $OUT/phase_split.txt 2>&1; cat $OUT/phase_split.txt; }
echo "== done"
