#!/bin/bash
# Sweep of environment-selected kernel builds under the full pipeline: one line per setting.
# usage: gpurun --timeout 1200 -- 'bash tools/gpu_sweep.sh tag "VAR1=a VAR2=b" "VAR1=c" ...'   (about 25 s per point)
TAG=${1:-sweep}
shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
i=0
for setting in "$@"; do
  i=$((i+1))
  name=p$i
  flags=""
  envs=""
  for tok in $setting; do
    case $tok in *=*) envs="$envs $tok";; *) flags="$flags $tok";; esac
  done
  env $envs timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --quick --no-parity --also "" $flags > $OUT/$name.json 2> $OUT/$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/$name.json").read().strip().splitlines()[-1])
    s = d["stages"]
    print("[$setting]", "value", round(d["value"]), "render", round(s["render_ms_per_view"] * 1e3, 1), "add", round(s["add_ms_per_view"] * 1e3, 1),
          "add_serial", round(s["add_ms_per_view_serial"] * 1e3, 1), "scatter", round(s["scatter_kernel_ms"] * 1e3, 1), "eager", round(s["api_eager_views_per_s"]), flush=True)
except Exception as e:
    print("[$setting] failed", e, open("$OUT/$name.err").read()[-300:], flush=True)
PY
done | tee $OUT/summary.txt
