#!/bin/bash
# A/B round: parity tests, then the default bench with the forward overlap on and off.
TAG=${1:-ab}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -15 $O/${TAG}_pytest.log
for ovl in 1 0; do
  UPS_OVERLAP_FWD=$ovl timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu --no-e2e > $O/${TAG}_bench_ovl$ovl.json 2> $O/${TAG}_bench_ovl$ovl.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/${TAG}_bench_ovl$ovl.json").read().strip().splitlines()[-1]); print("ovl=$ovl", round(d["value"]), d["ms_per_step"], d["per_call_ms"])
except Exception as e: print("ERR",e, open("$O/${TAG}_bench_ovl$ovl.err").read()[-1500:])
PY
done
