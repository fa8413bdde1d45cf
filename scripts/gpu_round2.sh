#!/bin/bash
# Round-2 one-GPU measurement round: parity tests, bench (all workloads), K4 variants, sweep (with K=25), TPS backward,
# reference arm, ncu launch list and full capture of the path kernels.   Usage (under gpurun): bash scripts/gpu_round2.sh <tag>
TAG=${1:-r02p}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > $O/${TAG}_clocks.csv &
SMI=$!
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -4 $O/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -2 $O/${TAG}_smoke.log
timeout 600 python bench.py --steps 100 --warmup 5 > $O/${TAG}_bench_cub.json 2> $O/${TAG}_bench_cub.err
timeout 600 python bench.py --steps 50 --warmup 5 --workload deepfashion --no-cpu --no-n4 --no-scale-workloads > $O/${TAG}_bench_deepfashion.json 2> $O/${TAG}_bench_df.err
timeout 600 python bench.py --steps 50 --warmup 5 --workload pennaction --no-cpu --no-n4 --no-scale-workloads > $O/${TAG}_bench_pennaction.json 2> $O/${TAG}_bench_penn.err
timeout 600 python bench.py --steps 50 --warmup 5 --tps-bwd --no-cpu --no-e2e --no-n4 --no-scale-workloads > $O/${TAG}_bench_tpsbwd.json 2> $O/${TAG}_bench_tpsbwd.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err
kill $SMI
python - <<PY
import json
for n in ("cub","deepfashion","pennaction","tpsbwd","reference"):
    try:
        d=json.loads(open("$O/${TAG}_bench_%s.json"%n).read().strip().splitlines()[-1])
        print(n, round(d["value"],1), d.get("ms_per_step"), d.get("step_roofline",{}).get("frac"), d.get("per_call_ms"), (d.get("e2e") or {}).get("value"), (d.get("e2e_uint8_views") or {}).get("value"))
    except Exception as e: print(n, "ERR", e)
PY
timeout 900 python scripts/sweep.py --out $O/${TAG}_sweep.json > $O/${TAG}_sweep.log 2>&1; tail -26 $O/${TAG}_sweep.log
timeout 300 python scripts/bench_inject_conv.py --tag ${TAG} > $O/${TAG}_inject_conv.log 2>&1
# launch list: same command as the bench, short
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-n4 --no-scale-workloads > $O/${TAG}_ncu_launches.log 2>&1
# full capture of each path kernel (one launch each, after warm-up): K13, K2, K4, K5 (+ K6 with --tps-bwd)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'step_|tps_warp' -s 15 -c 5 \
    -o $O/${TAG}_prof -f python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-n4 --no-scale-workloads --tps-bwd > $O/${TAG}_ncu_full.log 2>&1
ls -la $O | grep ${TAG} | tail -20
