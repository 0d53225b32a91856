for combo in "8 3" "4 3" "4 2" "6 2" "3 2" "2 3" "5 2"; do set -- $combo
  echo "== raster CTAs/SM $1, scatter CTAs/SM $2"
  SMESH_RASTER_CTAS=$1 SMESH_PAIR_CTAS=$2 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', round(d['value']), 'scatter_ms', round(d['stages']['scatter_kernel_ms'],4))"
done
