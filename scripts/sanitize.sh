#!/bin/bash
# compute-sanitizer passes over the tcgen05 / TMA / mbarrier kernels: memcheck (out-of-bounds, misaligned,
# invalid shared/global accesses) and racecheck (shared-memory hazards) on a small, representative subset of the GPU tests.
# Output: gpurun_out/sanitizer_{memcheck,racecheck}.log  (summarised under profiles/ by hand)
mkdir -p gpurun_out
SEL_OPS='(test_conv2d_tensor_path_vs_oracle and tf32 and (n4_c64_16x16_k64 or n3_c64 or n4_c64_16x16_k128_f3s2 or n2_c256 or n1_c32)) or (test_wgrad_haloed_tile_vs_oracle and (n3_c64 or n2_c32 or n2_c128 or n5_c64)) or test_pointwise_conv_with_few_filters or test_post_activation_block_tail_is_one_pass or test_residual_add_is_absorbed'
SEL_EPI='test_fused_epilogue_and_statistics and tf32 and (n4_c64 or n3_c64 or n2_c64)'
for tool in memcheck racecheck; do
  echo "== $tool" > gpurun_out/sanitizer_$tool.log
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 20 \
    python -m pytest tests/test_gpu_ops.py tests/test_gpu_epilogue.py tests/test_gpu_head.py -q -x -p no:cacheprovider \
      -k "($SEL_OPS) or ($SEL_EPI) or test_batch_norm_golden or test_relu_golden or log_softmax" >> gpurun_out/sanitizer_$tool.log 2>&1
  echo "exit code $?" >> gpurun_out/sanitizer_$tool.log
  tail -15 gpurun_out/sanitizer_$tool.log
done
