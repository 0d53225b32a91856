#!/bin/bash
# Multi-GPU bench lines of one box.  usage: gpurun --gpus N --timeout 1500 -- 'bash tools/gpu_multi.sh tag N'
TAG=${1:-multi}
N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
lscpu | grep -E "NUMA|Socket|Model name|^CPU\(s\)" > $OUT/cpu.txt
run() {  # name, extra args
  name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
    bench.py --gpus $N --steps 20 --warmup 3 "$@" > $OUT/$name.json 2> $OUT/$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/$name.json").read().strip().splitlines()[-1])
    s = d["stages"]
    print("$name N=$N", d["scaling"], "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"]),
          "h2d/gpu", round(d["e2e"]["h2d_gbs_plain_copy_per_gpu_min_over_ranks"], 1), "ceiling", round(d["e2e"]["h2d_ceiling_views_per_s"]),
          "allreduce ms in-region", round(s["allreduce_ms"], 3), "cold", round(s.get("allreduce_ms_cold_first_call", 0), 3), "warm", round(s.get("allreduce_ms_warm", 0), 3),
          "bus GB/s", round(s.get("allreduce_bus_gbs_warm", 0)), "parity", d["parity"].get("multi_gpu_vs_one_gpu_max_rel_err") if d["parity"] else None,
          "affinity", d["e2e"]["cpu_affinity"], flush=True)
except Exception as e:
    print("$name failed", e, open("$OUT/$name.err").read()[-600:], flush=True)
PY
}
run weak_cfg3 --quick
run strong_cfg3 --scaling strong --quick
run weak_cfg5 --config cfg5 --quick
