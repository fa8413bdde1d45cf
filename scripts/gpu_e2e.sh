#!/bin/bash
# parity tests, then the default bench including both e2e legs (fp32 and uint8 views).
TAG=${1:-e2e}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -6 $O/${TAG}_pytest.log
timeout 400 python bench.py --steps 100 --warmup 5 --no-cpu > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
python - <<PY
import json
try:
    d=json.loads(open("$O/${TAG}_bench.json").read().strip().splitlines()[-1]); print(round(d["value"]), d["ms_per_step"], d["per_call_ms"]); print("e2e", d["e2e"]); print("e2e_u8", d["e2e_uint8_views"])
except Exception as e: print("ERR",e, open("$O/${TAG}_bench.err").read()[-2500:])
PY
