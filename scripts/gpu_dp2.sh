#!/bin/bash
# 2-GPU data-parallel experiments: where does the scaling loss come from?
O=gpurun_out; TAG=${1:-dp2}; mkdir -p $O
run() { name=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus 2 --steps 60 --warmup 5 --no-e2e $EXTRA > $O/${TAG}_$name.json 2> $O/${TAG}_$name.err; python - <<PY
import json
try:
    d=json.loads(open("$O/${TAG}_$name.json").read().strip().splitlines()[-1]); print("$name", round(d["value"]), round(d["ms_per_step"],4), d["per_call_ms"])
except Exception as e: print("$name ERR", e, open("$O/${TAG}_$name.err").read()[-800:])
PY
}
EXTRA="" run default A=1
EXTRA="--grad-mb 0.001" run nocomm A=1
EXTRA="" run ctas8 NCCL_MAX_CTAS=8
EXTRA="" run ctas4 NCCL_MAX_CTAS=4
EXTRA="--bucket-mb 256" run onebucket A=1
EXTRA="--bucket-mb 256" run onebucket_ctas8 NCCL_MAX_CTAS=8
