#!/bin/bash
mkdir -p gpurun_out
echo "== phase split"; timeout 600 python scripts/phase_split.py > gpurun_out/phase_split.txt 2>&1; cat gpurun_out/phase_split.txt | tail -40
echo "== ncu full k_fused"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 4 -c 1 -f -o gpurun_out/prof_fused \
      python bench.py --envs 256 --steps 4 --warmup 3 --no-cpu-baseline --e2e-envs 16 --e2e-steps 2 --no-by-depth > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
