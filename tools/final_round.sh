#!/bin/bash
# Round-end evidence run (through gpurun): the whole -m gpu suite, the default bench line, the
# attention kernels alone, the profile set of tools/profile_round.sh and one ncu --set full capture
# of each attention kernel.  Small text / JSON files come back in gpurun_out/<tag>_*.
set -u
cd "$(dirname "$0")/.."
TAG=${1:-final}
mkdir -p gpurun_out
echo "== gpu tests"
timeout -s KILL 1500 python -m pytest tests -q -m gpu > gpurun_out/${TAG}_gputests.log 2>&1
echo "rc=$?"; tail -15 gpurun_out/${TAG}_gputests.log
cp gpurun_out/test_metrics.jsonl gpurun_out/${TAG}_test_metrics.jsonl 2>/dev/null
echo "== smoke"
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench"
timeout -s KILL 1500 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
echo "rc=$?"; cut -c1-400 gpurun_out/${TAG}_bench_n1.json; tail -2 gpurun_out/${TAG}_bench_n1.err
echo "== attention alone"
timeout -s KILL 300 python tools/attn_bench.py > gpurun_out/${TAG}_attn.json 2> gpurun_out/${TAG}_attn.err
cat gpurun_out/${TAG}_attn.json
echo "== profile set"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size,sm__warps_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed
python tools/update_kernels_bench.py > gpurun_out/${TAG}_update_kernels.log 2>&1
cp gpurun_out/update_kernels.json gpurun_out/${TAG}_update_kernels_hbm.json 2>/dev/null
timeout -s KILL 600 ncu --profile-from-start off --clock-control none --metrics $M -f -o /tmp/${TAG}_kernels \
  python tools/profile_kernels.py > gpurun_out/${TAG}_kernels_ncu.log 2>&1
python tools/ncu_summary.py table /tmp/${TAG}_kernels.ncu-rep > gpurun_out/${TAG}_kernels_ncu.txt 2>&1
timeout -s KILL 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file /tmp/${TAG}_launches.csv python tools/profile_kernels.py --step > gpurun_out/${TAG}_launches.log 2>&1
python tools/ncu_summary.py launches /tmp/${TAG}_launches.csv > gpurun_out/${TAG}_launches.txt 2>&1
head -14 gpurun_out/${TAG}_launches.txt
[ "${RVLM_NCU_FULL:-1}" = 1 ] || exit 0
echo "== ncu --set full: attention forward / backward"
for k in attn_fwd3 attn_bwd4; do
  timeout -s KILL 400 ncu --set full --import-source on --clock-control none -k regex:$k -c 1 -f \
    -o gpurun_out/${TAG}_$k python tools/attn_bench.py --reps 1 > gpurun_out/${TAG}_${k}_ncu.log 2>&1
  python tools/ncu_summary.py report gpurun_out/${TAG}_$k.ncu-rep > gpurun_out/${TAG}_${k}_ncu.txt 2>&1
  head -18 gpurun_out/${TAG}_${k}_ncu.txt
done
ls -la gpurun_out | tail -30
