#!/bin/bash
# One gpurun call: GPU tests, both bench arms, ncu launch list and one full capture of the scatter kernel.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh [tag]'
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/nvidia_smi.csv 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1
tail -5 $OUT/pytest_gpu.log
( time timeout 600 python bench.py ) > $OUT/bench_ours.json 2> $OUT/bench_ours.err
tail -c 3000 $OUT/bench_ours.json
( time SMESH_REF_BUDGET_S=${REF_BUDGET:-60} timeout 600 python bench.py --impl reference --steps 20 --warmup 1 ) > $OUT/bench_ref.json 2> $OUT/bench_ref.err
tail -c 1500 $OUT/bench_ref.json
# launch list: eager (no graph) so every kernel is its own launch; shares, not absolutes
SMESH_PROFILE_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > $OUT/launches_bench.log 2>&1
# full capture of the dominant kernel (scatter) on a few eager views of cfg3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scatter -s 4 -c 2 -o $OUT/scatter_full \
  python tools/prof_driver.py cfg3 4 > $OUT/scatter_full.log 2>&1
ls -la $OUT
