#!/bin/bash
# One gpurun call: smoke, GPU parity tests, bench, ncu launch list + full capture of k_fused.
# Usage (from the repo root on the GPU box): bash scripts/gpu_check.sh [quick|full]
MODE=${1:-full}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; grep -E "^(E   |FAILED|ERROR)|passed|failed" $OUT/pytest_gpu.log | cut -c1-300 | tail -12
[ -x scripts/microbench_atoms ] && { echo "== microbench atoms"; timeout 120 scripts/microbench_atoms | tee $OUT/microbench_atoms.log; }
echo "== bench"; timeout 900 python bench.py 2>$OUT/bench.err | tee $OUT/bench.json; tail -3 $OUT/bench.err
echo "== bench env8"; timeout 600 python bench.py --workload env8 --no-cpu-baseline 2>>$OUT/bench.err | tee $OUT/bench_env8.json
if [ "$MODE" = "full" ]; then
  echo "== ncu launch list"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches.csv \
      python bench.py --envs 256 --steps 4 --warmup 3 --no-cpu-baseline --e2e-envs 16 --e2e-steps 2 --no-by-depth --no-small-batch > $OUT/ncu_bench.log 2>&1
  grep -E "k_fused|k_cells|k_reset" $OUT/launches.csv | tail -12
  echo "== ncu full k_fused"
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 4 -c 2 -f -o $OUT/prof_fused \
      python bench.py --envs 256 --steps 4 --warmup 3 --no-cpu-baseline --e2e-envs 16 --e2e-steps 2 --no-by-depth --no-small-batch > $OUT/ncu_full.log 2>&1
  ls -la $OUT/*.ncu-rep
  [ -f ws-mgmap_b200/lib/libwsmg_phaseskip.so ] && { echo "== phase split"; timeout 600 python scripts/phase_split.py > $OUT/phase_split.txt 2>&1; tail -3 $OUT/phase_split.txt; }
  echo "== racecheck (small)"
  timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -q -x -k golden_trajectory 2>&1 | tail -8 | tee $OUT/racecheck.log
fi
echo "== done"
