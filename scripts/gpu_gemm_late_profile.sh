#!/bin/bash
# --set full captures of late-stage 1x1 GEMMs (block 17: 24x24, block 24: 12x12) in one batch-256 forward
mkdir -p gpurun_out
python scripts/profile_fwd.py 256 2 > gpurun_out/plain_fwd.log 2>&1 || { echo plain run failed; tail -5 gpurun_out/plain_fwd.log; exit 1; }
ncu --set full --import-source on --clock-control none -k 'regex:^pw_gemm_tc_kernel' -s 95 -c 2 -o gpurun_out/full_gemm_b17 -f python scripts/profile_fwd.py 256 2 > gpurun_out/ncu_full_gemm17.log 2>&1
echo "full gemm b17 exit=$?"
ncu --set full --import-source on --clock-control none -k 'regex:^pw_gemm_tc_kernel' -s 109 -c 2 -o gpurun_out/full_gemm_b24 -f python scripts/profile_fwd.py 256 2 > gpurun_out/ncu_full_gemm24.log 2>&1
echo "full gemm b24 exit=$?"
