#!/bin/bash
# compute-sanitizer over the parity shapes of the step and TPS kernels.
# Usage (under gpurun): bash scripts/gpu_sanitize.sh <tag>
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
SEL='test_config1_cub_b8 or test_fused_shapes or test_views_grad_tps_backward or test_decode_bwd_tensor_core_path or test_fused_forward_launch or test_unfused_fallback_shapes'
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 97 --print-limit 20 \
      python -m pytest tests/test_gpu_step.py tests/test_gpu_tps.py tests/test_gpu_inject_conv.py tests/test_gpu_dp.py -m gpu -x -q \
      -k "$SEL or test_warp or test_against_reference or test_padded_part_counts or test_parts_conv2d_backward_vs_oracle or test_standin or test_views_grad_without" \
      > $O/${TAG}_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?" >> $O/${TAG}_sanitizer_$tool.log
  tail -6 $O/${TAG}_sanitizer_$tool.log
done
