#!/bin/bash
# N=8: where the main gradient bucket is launched (before K4 = beside K4 and K5; after K4 = beside K5 only) x CTA count.
# Usage (under gpurun --gpus 8): bash scripts/gpu_dp8_variants.sh <tag>
TAG=${1:-r02m}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
for cfg in "0 16" "0 32" "1 16" "1 64"; do
  set -- $cfg
  UPS_DP_MAIN_AFTER_K4=$1 timeout 120 $TR bench.py --gpus 8 --steps 30 --warmup 5 --allreduce-ctas $2 --no-scale-workloads --no-e2e \
      > gpurun_out/${TAG}_bench_n8_a$1_c$2.json 2> gpurun_out/${TAG}_err.txt
  python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench_n8_a$1_c$2.json').read().strip().splitlines()[-1])
print('after_k4=$1 ctas=$2', d['value'], d['ms_per_step'], d['per_call_ms'])"
done
