#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
EXP_E=62 timeout 300 python scripts/exp/occupancy_exp.py 2>&1 | tee gpurun_out/occ_exp.txt
timeout 120 scripts/exp/microbench_tma 2>&1 | tee gpurun_out/microbench_tma.txt
{ lscpu | head -25; numactl -H; nvidia-smi topo -m; python -c "
import torch, os
import pynvml; pynvml.nvmlInit(); h=pynvml.nvmlDeviceGetHandleByIndex(0); b=pynvml.nvmlDeviceGetPciInfo(h).busId; print('busid', b)
print(os.sched_getaffinity(0), os.cpu_count())
b=b.decode() if isinstance(b,bytes) else b
b=b.lower()
if len(b.split(':')[0])==8: b=b[4:]
print(b, open('/sys/bus/pci/devices/%s/numa_node'%b).read())
"; } > gpurun_out/lscpu.txt 2>&1
