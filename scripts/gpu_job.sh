#!/bin/bash
# One gpurun job: GPU test suite, smoke, the bench line, optionally ncu passes.  Usage (from the repo root on the GPU box):
#   bash scripts/gpu_job.sh [tests] [smoke] [bench] [ncu_list] [ncu_full <kernel-regex> <name>] ...
# Everything is written under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
while [ $# -gt 0 ]; do
  case "$1" in
    tests)   timeout 1500 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS} > gpurun_out/test_all.log 2>&1; echo "tests rc=$?" ;;
    tests_s) timeout 1500 python -m pytest tests -m gpu -q -s ${PYTEST_ARGS} > gpurun_out/test_all.log 2>&1; echo "tests rc=$?" ;;
    smoke)   timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" ;;
    bench)   DFV_BENCH_DUMP=gpurun_out/infer_launches.json timeout 900 python bench.py --steps ${STEPS:-20} --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" ;;
    bench_ref) timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench_ref rc=$?" ;;
    ncu_list) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
                python scripts/profile_fwd.py 256 2 > gpurun_out/ncu.log 2>&1; echo "ncu_list rc=$?" ;;
    ncu_traffic) timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 600 --csv \
                --log-file gpurun_out/traffic.csv python scripts/profile_fwd.py 256 2 > gpurun_out/ncu_traffic.log 2>&1; echo "ncu_traffic rc=$?" ;;
    ncu_full) shift; rx="$1"; shift; nm="$1"; shift; skip="$1"
              timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$rx" -s ${skip:-0} -c 1 -f -o gpurun_out/full_$nm \
                python scripts/profile_fwd.py 256 2 > gpurun_out/ncu_full_$nm.log 2>&1; echo "ncu_full $nm rc=$?" ;;
    *) echo "running: $1"; timeout 1500 bash -c "$1"; echo "rc=$?" ;;
  esac
  shift
done
tail -5 gpurun_out/test_all.log 2>/dev/null
