#!/bin/bash
mkdir -p gpurun_out
K='regex:^(stem_kernel|dwconv_kernel|se_gate_kernel|pw_gemm_tc_kernel|hybrid_attention_kernel|mlp_head_kernel)'
python scripts/profile_fwd.py 256 2 > gpurun_out/plain_fwd.log 2>&1 || { echo plain run failed; tail -5 gpurun_out/plain_fwd.log; exit 1; }
# all kernels of the 2nd forward, light sections
ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section WarpStateStats --section LaunchStats --section SchedulerStats \
    --clock-control none -k "$K" -s 130 -c 130 -o gpurun_out/fwd_sections -f python scripts/profile_fwd.py 256 2 > gpurun_out/ncu_sections.log 2>&1
echo "sections exit=$?"
# full set + source for: dw b3 (k3), project b3, expand b4 | dw b17 (k5) , project b17
ncu --set full --import-source on --clock-control none -k "$K" -s 142 -c 4 -o gpurun_out/full_s2 -f python scripts/profile_fwd.py 256 2 > gpurun_out/ncu_full1.log 2>&1
echo "full1 exit=$?"
ncu --set full --import-source on --clock-control none -k "$K" -s 198 -c 3 -o gpurun_out/full_s5 -f python scripts/profile_fwd.py 256 2 > gpurun_out/ncu_full2.log 2>&1
echo "full2 exit=$?"
ls -la gpurun_out/*.ncu-rep
