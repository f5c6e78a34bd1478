#!/bin/bash
# Evidence for profiles/ (run on the GPU box through gpurun; only SMALL text files come back).
#   1. HBM roofline of the standalone update kernels (CUDA events, no profiler)
#   2. per-kernel ncu metrics (time, DRAM bytes, tensor pipe, issue slots) of every kernel of the library
#   3. ncu launch list of one full-size 10-iteration PGD call
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size,sm__warps_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed
python tools/update_kernels_bench.py > gpurun_out/r02_update_kernels.log 2>&1
cp gpurun_out/update_kernels.json gpurun_out/r02_update_kernels_hbm.json
timeout 600 ncu --profile-from-start off --clock-control none --metrics $M -f -o /tmp/r02_kernels \
  python tools/profile_kernels.py > gpurun_out/r02_kernels_ncu.log 2>&1
python tools/ncu_summary.py table /tmp/r02_kernels.ncu-rep > gpurun_out/r02_kernels_ncu.txt 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file /tmp/r02_launches.csv python tools/profile_kernels.py --step > gpurun_out/r02_launches.log 2>&1
python tools/ncu_summary.py launches /tmp/r02_launches.csv > gpurun_out/r02_launches.txt 2>&1
tail -3 gpurun_out/r02_kernels_ncu.log gpurun_out/r02_launches.log
wc -l gpurun_out/r02_kernels_ncu.txt gpurun_out/r02_launches.txt
