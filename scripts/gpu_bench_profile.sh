#!/bin/bash
# bench + ncu launch list (per-launch device time) for the same command.
mkdir -p gpurun_out
K='regex:^(stem_kernel|dwconv_kernel|se_gate_kernel|pw_gemm_tc_kernel|heat_raw_kernel|heat_norm_kernel|hybrid_attention_kernel|mlp_head_kernel)'
python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; tail -c 3000 gpurun_out/bench.json
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 396 -c 264 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "ncu exit=$?"
