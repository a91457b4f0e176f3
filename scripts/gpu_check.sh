#!/bin/bash
# Runs the GPU test groups in separate processes (a sticky CUDA fault in one group must not
# mask the others) and collects logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
run() {
  name=$1; shift
  timeout 600 python -m pytest "$@" -q --timeout 300 -p no:cacheprovider > gpurun_out/test_$name.log 2>&1
  echo "$name exit=$? $(tail -1 gpurun_out/test_$name.log)"
}
run stem tests/test_gpu_ops.py -k "stem"
run dwconv tests/test_gpu_ops.py -k "dwconv"
run se tests/test_gpu_ops.py -k "se_gate"
run gemm_all tests/test_gpu_ops.py -k "pw_gemm_all"
run gemm_big tests/test_gpu_ops.py -k "tcgen05_matches or rejects"
run attention tests/test_gpu_ops.py -k "heatmap or hybrid or mlp_head"
run loss tests/test_gpu_ops.py -k "combined_loss"
run model_fp32 tests/test_gpu_model.py -k "fp32 or state_dict"
run model_bf16 tests/test_gpu_model.py -k "bf16"
