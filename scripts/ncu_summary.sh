#!/bin/bash
# Turn gpurun_out/{prof_fused.ncu-rep,launches.csv,bench*.json} into text summaries under profiles/.
# Usage: bash scripts/ncu_summary.sh <tag>     (e.g. r01_v4)
TAG=${1:?tag}
OUT=profiles
cd "$(dirname "$0")/.."
ncu -i gpurun_out/prof_fused.ncu-rep --page raw --csv 2>/dev/null > /tmp/raw_$TAG.csv
python - "$TAG" <<'PY'
import csv, sys, json
tag = sys.argv[1]
rows = list(csv.reader(open(f"/tmp/raw_{tag}.csv")))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed_op_shared_atom.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
with open(f"profiles/{tag}_k_fused_ncu_full.txt", "w") as f:
    f.write(f"# ncu --set full --clock-control none -k regex:k_fused (bench.py --envs 256), {len(rows) - 2} launches captured\n")
    for li, r in enumerate(rows[2:]):
        f.write(f"## launch {li}: {r[hdr.index('Kernel Name')]}\n")
        for k in keys:
            if k in hdr:
                i = hdr.index(k)
                f.write(f"{k:78s} {r[i]:>18s} {units[i]}\n")
        rd = float(r[hdr.index('dram__bytes_read.sum')]); wr = float(r[hdr.index('dram__bytes_write.sum')])
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
        rd *= scale[units[hdr.index('dram__bytes_read.sum')]]; wr *= scale[units[hdr.index('dram__bytes_write.sum')]]
        mult = 1
        if li == 0:
            json.dump({"k_fused_dram_bytes_per_env": (rd + wr) * mult / 256, "source": f"profiles/{tag}_k_fused_ncu_full.txt",
                       "how": "ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum of one k_fused launch over 256 envs"},
                      open("profiles/ncu_traffic.json", "w"))
        f.write(f"dram traffic per launch (read+write)                                           {(rd + wr) * mult:18.0f} byte  (256 envs: {(rd + wr) * mult / 256 / 1e6:.2f} MB/env; algorithmic 20.79 MB/env)\n")
PY
ncu -i gpurun_out/prof_fused.ncu-rep --page source --csv --print-source cuda,sass 2>/dev/null > /tmp/src_$TAG.csv
{ echo "# executed-instruction / stall-sample breakdown of k_fused by SASS segment (scripts/ncu_segments.py)"; python scripts/ncu_segments.py /tmp/src_$TAG.csv 4096 0.012;
  echo; echo "# per phase (ranges of SASS between CTA-wide barriers, scripts/ncu_barrier_ranges.py)"; python scripts/ncu_barrier_ranges.py /tmp/src_$TAG.csv 4096;
  echo; echo "# LSU pipe: shared-memory wavefronts and global tag requests per instruction (scripts/ncu_wavefronts.py)"; python scripts/ncu_wavefronts.py /tmp/src_$TAG.csv 4096 24; } > $OUT/${TAG}_k_fused_source_breakdown.txt
{ echo "# ncu --metrics gpu__time_duration.sum --clock-control none (bench.py --envs 256 --steps 4 --warmup 3): per-launch device time, ns"; grep -E "k_fused|k_cells|k_reset" gpurun_out/launches.csv | awk -F'","' '{print $5, $(NF)}' | tr -d '"' | tail -24; } > $OUT/${TAG}_launches.txt
cp gpurun_out/bench.json $OUT/${TAG}_bench_sweep1024.json 2>/dev/null
cp gpurun_out/bench_env8.json $OUT/${TAG}_bench_env8.json 2>/dev/null
cp gpurun_out/microbench_atoms.log $OUT/${TAG}_microbench_atoms.txt 2>/dev/null
cp gpurun_out/racecheck.log $OUT/${TAG}_racecheck.txt 2>/dev/null
ls -la $OUT | tail -12
cp gpurun_out/phase_split.txt $OUT/${TAG}_phase_split.txt 2>/dev/null
