#!/bin/bash
# quick iteration: GPU tests + a short bench.  usage: gpurun --timeout 900 -- 'bash tools/gpu_quick.sh tag [pytest -k expr]'
TAG=${1:-quick}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( time timeout 600 python -m pytest tests -m gpu -x -q ${2:+-k "$2"} ) > $OUT/pytest_gpu.log 2>&1
tail -15 $OUT/pytest_gpu.log
( timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline ) > $OUT/bench.json 2> $OUT/bench.err
tail -c 400 $OUT/bench.err
python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
    print("value", round(d["value"]), "views/s; stages", {k: round(v, 4) for k, v in d["stages"].items()}, "roofline", round(d["roofline"]["frac"], 3), "e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("bench failed", e)
PY
