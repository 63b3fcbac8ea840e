#!/bin/bash
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest gpu"; timeout 2400 python -m pytest tests/test_policy_dropin.py tests/test_gpu_dropin.py -x -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1; grep -E "^(E   |FAILED|ERROR)|passed|failed|change cell" gpurun_out/pytest_gpu.log | cut -c1-300 | tail -14
echo "== bench default"; time timeout 1500 python bench.py 2>gpurun_out/bench.err > gpurun_out/bench.json;  python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'ms/step',round(d['ms_per_step'],3),'k_fused',round(d['roofline']['kernel_ms'],3),'frac',round(d['roofline']['frac'],3))
print('e2e',d['e2e']['value'],d['e2e']['pcie_ceiling_gbs'],d['e2e']['pcie_ceiling_frames_per_s'],d['e2e']['frac_of_pcie_ceiling'])
for k in ('small_batch','torch_cuda_baseline','cuda_reference_flips','policy_forward','cpu_baseline'):
    print(k, json.dumps(d.get(k))[:600])
PY
tail -5 gpurun_out/bench.err | cut -c1-300
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 10 --warmup 2 2>/dev/null | cut -c1-400
echo "== host overhead"; timeout 300 python scripts/host_overhead.py 2>&1 | tail -4
