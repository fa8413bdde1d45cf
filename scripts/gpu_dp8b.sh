#!/bin/bash
O=gpurun_out; TAG=${1:-dp8b}; mkdir -p $O
run() { name=$1; shift; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus 8 --steps 40 --warmup 5 --no-e2e > $O/${TAG}_$name.json 2> $O/${TAG}_$name.err; python - <<PY
import json
try:
    d=json.loads(open("$O/${TAG}_$name.json").read().strip().splitlines()[-1]); print("$name", round(d["value"]), round(d["ms_per_step"],4), d["per_call_ms"])
except Exception as e: print("$name ERR", e, open("$O/${TAG}_$name.err").read()[-600:])
PY
}
run default NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL,TUNING
grep -E "NVLS|Channel|channels|Algo|algo|nChannels|Connected" $O/${TAG}_default.err | sort | uniq -c | sort -rn | head -30 > $O/${TAG}_nccl_info.txt
run nvls NCCL_ALGO=NVLS
run ctas8 NCCL_MAX_CTAS=8
run nvls_ctas8 NCCL_ALGO=NVLS NCCL_MAX_CTAS=8
run ring NCCL_ALGO=Ring
