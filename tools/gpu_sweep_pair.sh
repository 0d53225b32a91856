#!/bin/bash
# How many scatter CTAs / consumer warps / ring stages per SM give the most views/s WITH the rasterizer on the same SMs?
# SMESH_PAIR_LEAN=1: the 64-register build of the kernel (8 CTAs of <= 3 consumer warps fit an SM's register file).
# usage: gpurun --timeout 900 -- 'bash tools/gpu_sweep_pair.sh [tag]'   (about 25 s per point)
TAG=${1:-sweep_pair}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for lean in 0 1; do
for ctas in 2 3 4; do
  for nw in 3 4; do
    [ $lean = 1 ] && [ $nw = 4 ] && continue
    for stages in 2 3; do
      name=lean${lean}_ctas${ctas}_nw${nw}_st${stages}
      SMESH_PAIR_LEAN=$lean SMESH_PAIR_CTAS=$ctas SMESH_PAIR_NW=$nw SMESH_PAIR_STAGES=$stages timeout 200 python bench.py --steps 20 --warmup 3 \
        --no-cpu-baseline > $OUT/$name.json 2> $OUT/$name.err
      python - <<PY
import json
try:
    d = json.loads(open("$OUT/$name.json").read().strip().splitlines()[-1])
    print("$name", "value", round(d["value"]), "render", round(d["stages"]["render_ms_per_view"] * 1e3, 1), "add",
          round(d["stages"]["add_ms_per_view"] * 1e3, 1), "scatter", round(d["stages"]["scatter_kernel_ms"] * 1e3, 1), flush=True)
except Exception as e:
    print("$name failed", e, flush=True)
PY
    done
  done
done
done | tee $OUT/summary.txt
