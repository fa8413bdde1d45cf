#!/bin/bash
O=gpurun_out; TAG=${1:-ovl}; mkdir -p $O
timeout 200 python scripts/bench_stats.py 2>&1 | tail -4
run() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu --no-e2e $EXTRA > $O/${TAG}_$name.json 2> $O/${TAG}_$name.err; python - <<PY
import json
try:
    d=json.loads(open("$O/${TAG}_$name.json").read().strip().splitlines()[-1]); print("$name", round(d["value"]), round(d["ms_per_step"],4), d["per_call_ms"])
except Exception as e: print("$name ERR", e, open("$O/${TAG}_$name.err").read()[-600:])
PY
}
EXTRA="" run t64_on UPS_FUSED_CTAS_PER_SM=64
EXTRA="" run t96_on UPS_FUSED_CTAS_PER_SM=96
EXTRA="" run t128_on UPS_FUSED_CTAS_PER_SM=128
EXTRA="" run t128_off UPS_FUSED_CTAS_PER_SM=128 UPS_OVERLAP_FWD=0
EXTRA="--workload deepfashion" run df_t64 UPS_FUSED_CTAS_PER_SM=64
EXTRA="--workload deepfashion" run df_t128 UPS_FUSED_CTAS_PER_SM=128
EXTRA="--workload deepfashion" run df_t256 UPS_FUSED_CTAS_PER_SM=256
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
