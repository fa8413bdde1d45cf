#!/bin/bash
# One GPU-box round: parity tests, bench (all workloads, both K4 variants), sweep, stats kernels,
# ncu launch list and full captures.   Usage (under gpurun): bash scripts/gpu_round.sh <tag>
TAG=${1:-r01}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > $O/${TAG}_clocks.csv &
SMI=$!
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -5 $O/${TAG}_pytest.log
timeout 600 python bench.py --steps 100 --warmup 5 > $O/${TAG}_bench_auto.json 2> $O/${TAG}_bench_auto.err
timeout 600 python bench.py --steps 100 --warmup 5 --decode-bwd simt --no-cpu --no-e2e > $O/${TAG}_bench_simt.json 2> $O/${TAG}_bench_simt.err
timeout 600 python bench.py --steps 50 --warmup 5 --workload deepfashion --no-cpu --no-e2e > $O/${TAG}_bench_df.json 2> $O/${TAG}_bench_df.err
timeout 600 python bench.py --steps 50 --warmup 5 --workload pennaction --no-cpu --no-e2e > $O/${TAG}_bench_penn.json 2> $O/${TAG}_bench_penn.err
timeout 600 python bench.py --steps 50 --warmup 5 --tps-bwd --no-cpu --no-e2e > $O/${TAG}_bench_tpsbwd.json 2> $O/${TAG}_bench_tpsbwd.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err
kill $SMI
cat $O/${TAG}_bench_auto.json
timeout 900 python scripts/sweep.py --out $O/${TAG}_sweep.json > $O/${TAG}_sweep.log 2>&1; tail -22 $O/${TAG}_sweep.log
timeout 300 python scripts/bench_stats.py > $O/${TAG}_stats_bench.log 2>&1; cat $O/${TAG}_stats_bench.log
# launch list: same command as the bench, short
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $O/${TAG}_ncu_launches.log 2>&1
# full capture of each path kernel (one launch each, after warm-up)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'step_|tps_warp' -s 15 -c 5 \
    -o $O/${TAG}_prof -f python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $O/${TAG}_ncu_full.log 2>&1
ls -la $O | tail -30
