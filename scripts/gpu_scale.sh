#!/bin/bash
# The driver's scaling run on one 8-GPU box: bench.py at N = 1, 2, 4, 8 with the driver's arguments.
# Usage (under gpurun --gpus 8): bash scripts/gpu_scale.sh <tag>
TAG=${1:-r02r}
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/${TAG}_topo.txt 2>&1
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/${TAG}_scale_n1.json 2> $O/${TAG}_scale_n1.err
for N in 2 4 8; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$N \
      bench.py --gpus $N --steps 20 --warmup 5 > $O/${TAG}_scale_n$N.json 2> $O/${TAG}_scale_n$N.err
done
python - <<PY
import json
base=None
for N in (1,2,4,8):
    try:
        d=json.loads(open("$O/${TAG}_scale_n%d.json"%N).read().strip().splitlines()[-1])
    except Exception as e:
        print(N,"ERR",e); continue
    if N==1: base=d
    sw=d.get("scale_workloads",{})
    print(N, "value %.0f"%d["value"], "eff %.3f"%(d["value"]/N/base["value"]), "ms %.4f"%d["ms_per_step"], d["config"]["allreduce"]["transport"],
          "e2e %.0f (x%.2f)"%(d["e2e"]["value"], d["e2e"]["value"]/base["e2e"]["value"]),
          "u8 %.0f (x%.2f)"%(d["e2e_uint8_views"]["value"], d["e2e_uint8_views"]["value"]/base["e2e_uint8_views"]["value"]),
          {k:(round(v.get("value",0)), round(v.get("step_roofline_frac",0),3)) for k,v in sw.items()}, d["per_call_ms"])
PY
