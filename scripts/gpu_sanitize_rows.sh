#!/bin/bash
# compute-sanitizer over the `_rows` entry points (ragged [.,K] rows read in place): memcheck on the full selection of
# gpu_sanitize.sh, racecheck on the padded-part-count and config-1 tests.   Usage: bash scripts/gpu_sanitize_rows.sh <tag>
TAG=${1:-r02q}
O=gpurun_out
mkdir -p $O
SEL='test_config1_cub_b8 or test_fused_shapes or test_views_grad_tps_backward or test_decode_bwd_tensor_core_path or test_fused_forward_launch or test_unfused_fallback_shapes'
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 97 --print-limit 20 \
    python -m pytest tests/test_gpu_step.py tests/test_gpu_tps.py tests/test_gpu_inject_conv.py tests/test_gpu_dp.py -m gpu -x -q \
    -k "$SEL or test_warp or test_against_reference or test_padded_part_count or test_parts_conv2d_backward_vs_oracle or test_standin or test_views_grad_without" \
    > $O/${TAG}_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" >> $O/${TAG}_sanitizer_memcheck.log
tail -5 $O/${TAG}_sanitizer_memcheck.log
timeout 200 compute-sanitizer --tool racecheck --error-exitcode 97 --print-limit 20 \
    python -m pytest tests/test_gpu_step.py -m gpu -x -q -k "test_padded_part_count or test_config1_cub_b8" \
    > $O/${TAG}_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" >> $O/${TAG}_sanitizer_racecheck.log
tail -5 $O/${TAG}_sanitizer_racecheck.log
