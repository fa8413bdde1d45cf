#!/bin/bash
# N4 kernels: parity tests, microbench, ncu.   Usage (under gpurun): bash scripts/gpu_n4.sh <tag> [ncu]
TAG=${1:-r01d}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_inject_conv.py -m gpu -q 2>&1 | grep -E "AssertionError:|passed|failed|Error|FAILED" | head -40 > $O/${TAG}_pytest_n4.log; cat $O/${TAG}_pytest_n4.log
timeout 300 python scripts/bench_inject_conv.py --tag $TAG > $O/${TAG}_inject_conv.log 2>&1; cat $O/${TAG}_inject_conv.log | tail -20
if [ "$2" = "ncu" ]; then
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:"inject_conv|parts_conv" -c 8 -o $O/${TAG}_n4_prof -f python scripts/bench_inject_conv.py --no-library --once > $O/${TAG}_n4_ncu.log 2>&1; tail -3 $O/${TAG}_n4_ncu.log
fi
