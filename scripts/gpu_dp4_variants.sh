#!/bin/bash
# N=4: main-bucket launch position x CTA count.  Usage (under gpurun --gpus 4): bash scripts/gpu_dp4_variants.sh <tag>
TAG=${1:-r02s}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514"
for cfg in "0 32" "0 64" "1 32" "1 64"; do
  set -- $cfg
  UPS_DP_MAIN_AFTER_K4=$1 timeout 120 $TR bench.py --gpus 4 --steps 30 --warmup 5 --allreduce-ctas $2 --no-scale-workloads --no-e2e \
      > gpurun_out/${TAG}_bench_n4_a$1_c$2.json 2> gpurun_out/${TAG}_err.txt
  python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench_n4_a$1_c$2.json').read().strip().splitlines()[-1])
print('after_k4=$1 ctas=$2', d['value'], d['ms_per_step'], d['per_call_ms'])"
done
timeout 100 $TR scripts/bench_allreduce.py --ctas 16,32,64,128 2>/dev/null | tee gpurun_out/${TAG}_allreduce_n4.json
