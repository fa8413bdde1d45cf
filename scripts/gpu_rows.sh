#!/bin/bash
# In-place ragged rows check: full GPU suite, smoke, K=25 (contiguous) and K=16 bench lines.   Usage: bash scripts/gpu_rows.sh <tag>
TAG=${1:-r02q}
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -4 $O/${TAG}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -2 $O/${TAG}_smoke.log
timeout 200 python bench.py --steps 50 --warmup 5 --n-parts 25 --no-cpu --no-n4 --no-scale-workloads > $O/${TAG}_bench_cub_k25.json 2> $O/${TAG}_bench_k25.err
timeout 200 python bench.py --steps 50 --warmup 5 --no-cpu --no-n4 --no-scale-workloads > $O/${TAG}_bench_cub.json 2> $O/${TAG}_bench_cub.err
python - <<PY
import json
for n in ("cub_k25","cub"):
    try:
        d=json.loads(open("$O/${TAG}_bench_%s.json"%n).read().strip().splitlines()[-1])
        print(n, round(d["value"],1), d.get("ms_per_step"), d.get("gpu_launches"), d.get("per_call_ms"), (d.get("e2e") or {}).get("value"), (d.get("e2e_uint8_views") or {}).get("value"))
    except Exception as e: print(n, "ERR", e)
PY
