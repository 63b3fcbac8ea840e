cd $GRAFT_REPO_ROOT
ARGS="--shape b256c27 --envs 128 --steps 3 --warmup 3 --no-cpu-baseline --e2e-envs 8 --e2e-steps 2 --no-by-depth --no-small-batch"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 3 -c 1 -f -o gpurun_out/prof_c27 python bench.py $ARGS > gpurun_out/ncu_c27.log 2>&1
ls -la gpurun_out/prof_c27.ncu-rep
