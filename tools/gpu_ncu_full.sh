#!/bin/bash
# full ncu capture of kernels matching a regex on a few eager views. usage: gpurun -- 'bash tools/gpu_ncu_full.sh tag regex [cfg] [skip] [count]'
TAG=${1:-full}; RE=${2:-scatter}; CFG=${3:-cfg3}; SKIP=${4:-4}; CNT=${5:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SKIP -c $CNT -o $OUT/full_${CFG} -f \
  python tools/prof_driver.py $CFG 4 > $OUT/full_${CFG}.log 2>&1
tail -3 $OUT/full_${CFG}.log
ls -la $OUT
