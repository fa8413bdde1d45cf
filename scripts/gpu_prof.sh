#!/bin/bash
# Profiling round: parity tests, default bench, ncu launch list + full capture of every path kernel.
# Usage (under gpurun): bash scripts/gpu_prof.sh <tag>
TAG=${1:-r01b}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -5 $O/${TAG}_pytest.log
timeout 600 python bench.py --steps 100 --warmup 5 > $O/${TAG}_bench_cub.json 2> $O/${TAG}_bench_cub.err
cat $O/${TAG}_bench_cub.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $O/${TAG}_ncu_launches.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'step_|tps_warp' -s 15 -c 5 \
    -o $O/${TAG}_prof -f python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $O/${TAG}_ncu_full.log 2>&1
ls -la $O | tail -12
