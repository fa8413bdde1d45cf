#!/bin/bash
O=gpurun_out; TAG=${1:-dp8}; mkdir -p $O
for N in 8 4; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus $N --steps 60 --warmup 5 > $O/${TAG}_n$N.json 2> $O/${TAG}_n$N.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/${TAG}_n$N.json").read().strip().splitlines()[-1]); print($N, round(d["value"]), round(d["ms_per_step"],4), d["per_call_ms"], "e2e", round(d["e2e"]["value"]))
except Exception as e: print("$N ERR", e, open("$O/${TAG}_n$N.err").read()[-800:])
PY
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus 8 --steps 60 --warmup 5 --no-e2e --workload pennaction > $O/${TAG}_penn_n8.json 2> $O/${TAG}_penn_n8.err; tail -c 600 $O/${TAG}_penn_n8.json
