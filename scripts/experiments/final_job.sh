#!/bin/bash
# Last GPU call of the round: GPU suite, bench line (+ per-launch dumps) and smoke() on the final build.
mkdir -p gpurun_out
timeout 100 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_last.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/gpu_tests_last.log
DFV_BENCH_DUMP=gpurun_out/launches_last.json timeout 80 python bench.py > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err; echo "bench rc=$?"
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_last.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_last.log
