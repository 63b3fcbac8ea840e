cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ARGS="--envs 256 --steps 4 --warmup 3 --no-cpu-baseline --e2e-envs 16 --e2e-steps 2 --no-by-depth --no-small-batch"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches256.csv python bench.py $ARGS > gpurun_out/ncu_l.log 2>&1
grep -E "k_fused|k_cells|k_reset" gpurun_out/launches256.csv | awk -F'","' '{print $5, $(NF-6), $(NF)}' | tr -d '"' | cut -c1-200 | tail -16
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_cells -s 4 -c 1 -f -o gpurun_out/prof_cells python bench.py $ARGS > gpurun_out/ncu_c.log 2>&1
ls -la gpurun_out/prof_cells.ncu-rep
