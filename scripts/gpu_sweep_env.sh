#!/bin/bash
# Sweep env-selected kernel variants: usage  gpu_sweep_env.sh <tag> "VAR=a VAR2=b" "VAR=c" ...
TAG=$1; shift
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q -k "tps or step or softmax or parts" > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -4 $O/${TAG}_pytest.log
i=0
for cfg in "$@"; do
  i=$((i+1))
  env $cfg timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu --no-e2e > $O/${TAG}_bench_$i.json 2> $O/${TAG}_bench_$i.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/${TAG}_bench_$i.json").read().strip().splitlines()[-1]); print("$cfg", round(d["value"]), round(d["ms_per_step"],4), d["per_call_ms"])
except Exception as e: print("$cfg", "ERR",e, open("$O/${TAG}_bench_$i.err").read()[-800:])
PY
done
