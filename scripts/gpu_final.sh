#!/bin/bash
# Round-end GPU pass: full parity suite, smoke, bench (all workloads + reference arm), launch list, N4 bench + ncu.
# Usage (under gpurun): bash scripts/gpu_final.sh <tag>
TAG=${1:-r01d}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > $O/${TAG}_clocks.csv &
SMI=$!
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -4 $O/${TAG}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -1 $O/${TAG}_smoke.log
timeout 600 python bench.py --steps 100 --warmup 5 > $O/${TAG}_bench_cub.json 2> $O/${TAG}_bench_cub.err
timeout 300 python bench.py --steps 50 --warmup 5 --workload deepfashion --no-cpu --no-e2e > $O/${TAG}_bench_deepfashion.json 2> $O/${TAG}_bench_df.err
timeout 300 python bench.py --steps 50 --warmup 5 --workload pennaction --no-cpu --no-e2e > $O/${TAG}_bench_pennaction.json 2> $O/${TAG}_bench_penn.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err
kill $SMI
cat $O/${TAG}_bench_cub.json
timeout 300 python scripts/bench_inject_conv.py --tag $TAG > $O/${TAG}_inject_conv.log 2>&1; tail -12 $O/${TAG}_inject_conv.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-n4 > $O/${TAG}_ncu_launches.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"inject_conv|parts_conv" -c 8 -o $O/${TAG}_n4_prof -f \
    python scripts/bench_inject_conv.py --no-library --once > $O/${TAG}_n4_ncu.log 2>&1; tail -2 $O/${TAG}_n4_ncu.log
ls -la $O | tail -20
