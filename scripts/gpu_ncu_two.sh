#!/bin/bash
mkdir -p gpurun_out
K='regex:^(stem_kernel|dwconv_kernel|se_gate_kernel|pw_gemm_tc_kernel|hybrid_attention_kernel|mlp_head_kernel)'
python scripts/profile_fwd.py 256 2 > gpurun_out/plain_fwd.log 2>&1 || { echo plain run failed; tail -5 gpurun_out/plain_fwd.log; exit 1; }
# 2nd forward starts at filtered index 130; dw b3 = +12, proj b3 = +14, expand b4 = +15 ; stem = +0
ncu --set full --import-source on --clock-control none -k "$K" -s ${1:-142} -c ${2:-4} -o gpurun_out/full_a -f python scripts/profile_fwd.py 256 2 > gpurun_out/ncu_full_a.log 2>&1
echo "full_a exit=$?"
if [ -n "$3" ]; then
ncu --set full --import-source on --clock-control none -k "$K" -s $3 -c ${4:-1} -o gpurun_out/full_b -f python scripts/profile_fwd.py 256 2 > gpurun_out/ncu_full_b.log 2>&1
echo "full_b exit=$?"
fi
