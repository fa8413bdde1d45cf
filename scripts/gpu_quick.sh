#!/bin/bash
# Quick round: parity tests, then the default bench under a few K1 register-cap variants.
# Usage (under gpurun): bash scripts/gpu_quick.sh <tag> [minb values...]
TAG=${1:-q}; shift
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -6 $O/${TAG}_pytest.log
for mb in ${@:-8}; do
  UPS_TPS_MINB=$mb timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu --no-e2e > $O/${TAG}_bench_minb$mb.json 2> $O/${TAG}_bench_minb$mb.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/${TAG}_bench_minb$mb.json").read().strip().splitlines()[-1]); print("minb=$mb", round(d["value"]), d["ms_per_step"], d["per_call_ms"])
except Exception as e: print("ERR",e, open("$O/${TAG}_bench_minb$mb.err").read()[-1500:])
PY
done
UPS_OVERLAP_FWD=0 timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu --no-e2e > $O/${TAG}_bench_noovl.json 2> $O/${TAG}_bench_noovl.err
python - <<PY
import json
d=json.loads(open("$O/${TAG}_bench_noovl.json").read().strip().splitlines()[-1]); print("no overlap", round(d["value"]), d["ms_per_step"], d["per_call_ms"])
PY
