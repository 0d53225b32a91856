#!/bin/bash
# One gpurun call of round 2: GPU tests, the bench line, optionally launch list / full captures.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_r02.sh tag [tests|notests] [ncu]'
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/nvidia_smi.csv 2>&1
nproc > $OUT/nproc.txt; lscpu | head -20 >> $OUT/nproc.txt
if [ "${2:-tests}" = "tests" ]; then
  ( time timeout 1200 python -m pytest tests -m gpu -x -q --durations=15 ) > $OUT/pytest_gpu.log 2>&1
  tail -30 $OUT/pytest_gpu.log
fi
( time timeout 900 python bench.py ) > $OUT/bench_ours.json 2> $OUT/bench_ours.err
tail -c 600 $OUT/bench_ours.err
python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_ours.json").read().strip().splitlines()[-1])
    print("value", round(d["value"]), "views/s; ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"]))
    print("stages", {k: round(v, 4) for k, v in d["stages"].items()})
    print("roofline", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d["roofline"].items() if k.startswith("frac") or k in ("achieved", "input_only_frac")})
    print("parity", d["parity"])
    print("other", json.dumps(d.get("other_configs"), indent=0)[:1500])
    print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("bench failed", e)
PY
if [ "${3:-}" = "ncu" ]; then
  SMESH_PROFILE_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 3 --cycles 2 --no-graph --no-cpu-baseline --quick --also '' > $OUT/launches_bench.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:scatter_pair -s 4 -c 2 -o $OUT/add_full \
    python tools/prof_driver.py cfg3 4 > $OUT/add_full.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:raster_unit -s 4 -c 1 -o $OUT/raster_full \
    python tools/prof_driver.py cfg3 4 > $OUT/raster_full.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:get_stream -c 1 -o $OUT/get_full \
    python tools/prof_driver.py cfg3 4 > $OUT/get_full.log 2>&1
fi
ls -la $OUT
