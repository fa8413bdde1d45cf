#!/bin/bash
# Two-GPU check of the final tree: the DP tests that need two ranks and one short N=2 bench line.  Usage: bash scripts/gpu_final2.sh <tag>
TAG=${1:-r02v}
O=gpurun_out
mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_dp.py -m gpu -x -q > $O/${TAG}_pytest_dp2.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_dp2.log
tail -3 $O/${TAG}_pytest_dp2.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu --no-n4 > $O/${TAG}_bench_n2.json 2> $O/${TAG}_bench_n2.err
python - <<PY
import json
d=json.loads(open("$O/${TAG}_bench_n2.json").read().strip().splitlines()[-1])
print(d["n_gpus"], round(d["value"]), d["ms_per_step"], d.get("allreduce"), d["roofline"]["frac"], d["e2e"]["value"], d["e2e_uint8_views"]["value"], {k: v.get("value") for k, v in d.get("scale_workloads", {}).items()})
PY
