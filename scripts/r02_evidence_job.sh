#!/bin/bash
# Round-2 evidence job (one gpurun call): fp32 batch-1 diagnosis, GPU tests, smoke, both bench arms, ncu launch list / DRAM traffic
# of one forward, --set full captures of the 5x5 depthwise layer of block 7 and the two GEMMs of block 24.
mkdir -p gpurun_out
timeout 600 python scripts/diag_fp32_b1.py > gpurun_out/diag_fp32.log 2>&1; echo "diag rc=$?"
bash scripts/gpu_job.sh tests smoke bench bench_ref ncu_list ncu_traffic ncu_full dwconv_kernel dw7 7
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:pw_gemm_tc_kernel" -s 47 -c 2 -f -o gpurun_out/full_gemm_b24 \
  python scripts/profile_fwd.py 256 2 > gpurun_out/ncu_full_gemm_b24.log 2>&1; echo "ncu_full gemm_b24 rc=$?"
