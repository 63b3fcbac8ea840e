cd $GRAFT_REPO_ROOT
echo "== pytest spc=4"; WSMG_SLABS_PER_CTA=4 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -x -q 2>&1 | grep -E "^(E   |FAILED|ERROR)|passed|failed" | cut -c1-200 | tail -5
for s in 1 2 4 8; do echo "== bench spc=$s"; WSMG_SLABS_PER_CTA=$s timeout 600 python bench.py --no-cpu-baseline --no-by-depth --e2e-envs 16 --e2e-steps 2 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms_per_step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'small', d.get('small_batch', {}).get('ms_per_step'))"; done
