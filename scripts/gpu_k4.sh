#!/bin/bash
# K4 bring-up round: parity of the decode_bwd variants, then A/B bench.
TAG=${1:-k4}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_step.py -x -q -k "decode_bwd_tensor_core" > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -25 $O/${TAG}_pytest.log
if grep -q "pytest rc=0" $O/${TAG}_pytest.log; then
  timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu --no-e2e > $O/${TAG}_bench_tc.json 2> $O/${TAG}_bench_tc.err
  timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu --no-e2e --decode-bwd simt > $O/${TAG}_bench_simt.json 2> $O/${TAG}_bench_simt.err
  python - <<PY
import json
for n in ("tc","simt"):
    try:
        d=json.loads(open("$O/${TAG}_bench_%s.json"%n).read().strip().splitlines()[-1]); print(n, round(d["value"]), d["ms_per_step"], d["per_call_ms"])
    except Exception as e: print(n,"ERR",e, open("$O/${TAG}_bench_%s.err"%n).read()[-1500:])
PY
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'decode_bwd_tma' -s 3 -c 1 \
    -o $O/${TAG}_prof -f python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $O/${TAG}_ncu_full.log 2>&1
  timeout 600 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_all.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_all.log; tail -3 $O/${TAG}_pytest_all.log
fi
