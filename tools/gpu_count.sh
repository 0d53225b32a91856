#!/bin/bash
# count-stage variants: timings + one ncu metrics pass.  usage: gpurun --timeout 900 -- 'bash tools/gpu_count.sh tag'
TAG=${1:-count}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python tools/time_count.py cfg3 8 2>&1 | tee $OUT/time_count_cfg3.txt
timeout 300 python tools/time_count.py cfg5 8 2>&1 | tee $OUT/time_count_cfg5.txt
timeout 300 python tools/time_count.py cfg2 8 2>&1 | tee $OUT/time_count_cfg2.txt
NCU=1 timeout 600 ncu --metrics gpu__time_duration.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum,l1tex__t_requests_pipe_lsu_mem_global_op_red.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,dram__bytes_read.sum,sm__inst_executed_pipe_lsu.sum \
  --clock-control none -k regex:count_ --csv --log-file $OUT/ncu_count.csv python tools/time_count.py cfg3 2 > $OUT/ncu_count.log 2>&1
python - <<PY
import csv
rows = list(csv.reader(open("$OUT/ncu_count.csv")))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
idx = {n: i for i, n in enumerate(rows[hdr])}
out = {}
for r in rows[hdr + 1:]:
    if len(r) < len(idx): continue
    key = (r[idx["ID"]], r[idx["Kernel Name"]][:60])
    out.setdefault(key, {})[r[idx["Metric Name"]]] = r[idx["Metric Value"]]
for k, v in out.items():
    print(k, v)
PY
