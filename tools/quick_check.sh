#!/bin/bash
# Quick GPU check of a kernel change (run through gpurun): block-level parity tests, the attention
# kernels timed alone, a short bench line and the ncu launch list of one full-size PGD call.
# Everything lands in gpurun_out/<tag>_*; the tag is the first argument.
set -u
cd "$(dirname "$0")/.."
TAG=${1:-quick}
mkdir -p gpurun_out
echo "== block tests"
timeout -s KILL 900 python -m pytest tests/test_blocks_gpu.py -x -q > gpurun_out/${TAG}_blocks.log 2>&1
echo "rc=$?"; tail -5 gpurun_out/${TAG}_blocks.log
echo "== attention alone"
timeout -s KILL 300 python tools/attn_bench.py > gpurun_out/${TAG}_attn.json 2> gpurun_out/${TAG}_attn.err
RVLM_ATTN_FWD_ONE=0 timeout -s KILL 300 python tools/attn_bench.py >> gpurun_out/${TAG}_attn.json 2>> gpurun_out/${TAG}_attn.err
cat gpurun_out/${TAG}_attn.json; tail -3 gpurun_out/${TAG}_attn.err
echo "== engine tests"
timeout -s KILL 900 python -m pytest tests/test_engine_gpu.py tests/test_attacks_gpu.py -x -q > gpurun_out/${TAG}_engine.log 2>&1
echo "rc=$?"; tail -5 gpurun_out/${TAG}_engine.log
echo "== bench (short)"
timeout -s KILL 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-modes --no-torch-baseline \
  --no-train-step > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "rc=$?"; cut -c1-700 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
echo "== launch list"
timeout -s KILL 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file /tmp/${TAG}_launches.csv python tools/profile_kernels.py --step > gpurun_out/${TAG}_launches.log 2>&1
python tools/ncu_summary.py launches /tmp/${TAG}_launches.csv > gpurun_out/${TAG}_launches.txt 2>&1
head -16 gpurun_out/${TAG}_launches.txt
