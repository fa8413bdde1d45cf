#!/bin/bash
O=gpurun_out; TAG=${1:-dp8c}; mkdir -p $O
run() { name=$1; shift; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus 8 --steps 40 --warmup 5 --no-e2e > $O/${TAG}_$name.json 2> $O/${TAG}_$name.err; grep -o '^{"metric.*' $O/${TAG}_$name.json | tail -1 | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$name', round(d['value']), round(d['ms_per_step'],4), d['per_call_ms'])"; grep -h "AllReduce: .*Algo" $O/${TAG}_$name.json $O/${TAG}_$name.err | sort | uniq -c | head -3; }
run sum NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=COLL,TUNING
run sum_nvls NCCL_ALGO=NVLS
run sum_nvlstree NCCL_ALGO=NVLSTree
