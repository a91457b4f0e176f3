#!/bin/bash
mkdir -p gpurun_out
K='regex:^(stem_kernel|dwconv_kernel|se_gate_kernel|pw_gemm_tc_kernel|hybrid_attention_kernel)'
python scripts/profile_fwd.py 256 2 > gpurun_out/plain_fwd.log 2>&1 || { echo plain run failed; tail -5 gpurun_out/plain_fwd.log; exit 1; }
ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section WarpStateStats --section LaunchStats --section SchedulerStats \
    --clock-control none -k "$K" -s 129 -c 129 -o gpurun_out/fwd_sections -f python scripts/profile_fwd.py 256 2 > gpurun_out/ncu_sections.log 2>&1
echo "sections exit=$?"
ncu --set full --import-source on --clock-control none -k 'regex:^dwconv_kernel' -s 39 -c 1 -o gpurun_out/full_dw7 -f python scripts/profile_fwd.py 256 2 > gpurun_out/ncu_full1.log 2>&1
echo "full dw7 exit=$?"
ncu --set full --import-source on --clock-control none -k 'regex:^dwconv_kernel' -s 55 -c 1 -o gpurun_out/full_dw23 -f python scripts/profile_fwd.py 256 2 > gpurun_out/ncu_full1b.log 2>&1
echo "full dw23 exit=$?"
ncu --set full --import-source on --clock-control none -k 'regex:^pw_gemm_tc_kernel' -s 64 -c 2 -o gpurun_out/full_gemm12 -f python scripts/profile_fwd.py 256 2 > gpurun_out/ncu_full2.log 2>&1
echo "full gemm exit=$?"
ncu --set full --import-source on --clock-control none -k 'regex:^stem_kernel' -s 1 -c 1 -o gpurun_out/full_stem -f python scripts/profile_fwd.py 256 2 > gpurun_out/ncu_full3.log 2>&1
echo "full stem exit=$?"
ls -la gpurun_out/*.ncu-rep
