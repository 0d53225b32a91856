#!/bin/bash
# launch list of our kernels for a few eager views. usage: gpurun -- 'bash tools/gpu_launches.sh tag [cfg] [views]'
TAG=${1:-launches}; CFG=${2:-cfg3}; VIEWS=${3:-6}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'view_begin|raster_|resolve|scatter|count_|clear_kernel' \
  --csv --log-file $OUT/launches_$CFG.csv python tools/prof_driver.py $CFG $VIEWS > $OUT/launches_$CFG.log 2>&1
python tools/ncu_launches.py $OUT/launches_$CFG.csv
