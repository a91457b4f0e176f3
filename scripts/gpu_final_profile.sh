#!/bin/bash
# Round-end evidence: tests, both benches (with CPU baseline), ncu launch list of one forward, DRAM traffic of the
# dominant kernel's launches, one --set full capture of the top kernel.
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 python -m pytest "$@" -q --timeout 600 -p no:cacheprovider > gpurun_out/test_$name.log 2>&1; echo "$name exit=$? $(tail -1 gpurun_out/test_$name.log)"; }
run all tests -m gpu
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks.csv &
SMI=$!
DFV_BENCH_DUMP=gpurun_out/infer_launches.json python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"
python bench.py --mode train --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "bench train exit=$?"
kill $SMI
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench reference exit=$?"
python scripts/profile_fwd.py 256 2 > gpurun_out/plain_fwd.log 2>&1 || { echo plain run failed; tail -5 gpurun_out/plain_fwd.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/profile_fwd.py 256 2 > gpurun_out/ncu.log 2>&1
echo "launch list exit=$?"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k 'regex:^(dwconv_kernel|pw_gemm_tc_kernel|stem_tc_kernel)' -s 96 -c 96 --csv --log-file gpurun_out/traffic.csv python scripts/profile_fwd.py 256 2 > gpurun_out/ncu_traffic.log 2>&1
echo "traffic exit=$?"
ncu --set full --import-source on --clock-control none -k 'regex:^dwconv_kernel' -s 35 -c 1 -o gpurun_out/full_dw3 -f python scripts/profile_fwd.py 256 2 > gpurun_out/ncu_full_dw3.log 2>&1
echo "full dw3 exit=$?"
ncu --set full --import-source on --clock-control none -k 'regex:^dwconv_kernel' -s 39 -c 1 -o gpurun_out/full_dw7 -f python scripts/profile_fwd.py 256 2 > gpurun_out/ncu_full_dw7.log 2>&1
echo "full dw7 exit=$?"
ncu --set full --import-source on --clock-control none -k 'regex:^pw_gemm_tc_kernel' -s 67 -c 2 -o gpurun_out/full_gemm_b3 -f python scripts/profile_fwd.py 256 2 > gpurun_out/ncu_full_gemm.log 2>&1
echo "full gemm exit=$?"
ncu --set full --import-source on --clock-control none -k 'regex:^pw_gemm_tc_kernel' -s 109 -c 2 -o gpurun_out/full_gemm_b24 -f python scripts/profile_fwd.py 256 2 > gpurun_out/ncu_full_gemm24.log 2>&1
echo "full gemm b24 exit=$?"
