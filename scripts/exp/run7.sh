#!/bin/bash
mkdir -p gpurun_out
echo "== pytest policy"; timeout 1200 python -m pytest tests/test_policy_dropin.py -x -q -m gpu > gpurun_out/pytest_policy.log 2>&1; grep -E "^(E   |FAILED|ERROR)|passed|failed" gpurun_out/pytest_policy.log | cut -c1-300 | tail -8
echo "== bench --gpus 2"; time timeout 1500 python bench.py --gpus 2 2>gpurun_out/bench2.err > gpurun_out/bench_n2.json; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
print('N=2 value',round(d['value']),'ms/step',round(d['ms_per_step'],3),'k_fused',round(d['roofline']['kernel_ms'],3),'frac',round(d['roofline']['frac'],3), d['config']['workload'][:60])
print('weak',d.get('weak'))
print('e2e',d['e2e']['value'],d['e2e']['pcie_ceiling_gbs'],d['e2e']['pcie_ceiling_frames_per_s'],d['e2e']['frac_of_pcie_ceiling'], d['e2e']['numa_note'])
PY
tail -3 gpurun_out/bench2.err | cut -c1-300
echo "== bench --gpus 2 reference"; timeout 600 python bench.py --gpus 2 --impl reference --steps 5 --warmup 1 2>/dev/null | cut -c1-200
