#!/bin/bash
# N=8: the main bucket in one piece before K4 (16 / 24 CTAs) vs split: first half before K4 (16 CTAs), second half after
# K4 with more CTAs.   Usage (under gpurun --gpus 8): bash scripts/gpu_dp8_split.sh <tag>
TAG=${1:-r02t}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518"
for cfg in "0 16 0" "32 16 0" "64 16 0" "0 128 1"; do
  set -- $cfg
  UPS_DP_MAIN_AFTER_K4=$3 UPS_DP_SPLIT=$1 timeout 120 $TR bench.py --gpus 8 --steps 30 --warmup 5 --allreduce-ctas $2 --no-scale-workloads --no-e2e \
      > gpurun_out/${TAG}_bench_n8_s$1_c$2_a$3.json 2> gpurun_out/${TAG}_err.txt
  python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench_n8_s$1_c$2_a$3.json').read().strip().splitlines()[-1])
print('split_ctas=$1 ctas=$2 after_k4=$3', d['value'], d['ms_per_step'], d['per_call_ms'])"
done
