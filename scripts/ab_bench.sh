#!/bin/bash
# quick A/B job: op + model tests, then the bench line with the per-launch dump.  Output under gpurun_out/.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_edge_cases.py tests/test_gpu_side_apis.py -x -q > gpurun_out/test_gemm.log 2>&1; echo "tests rc=$?"
DFV_BENCH_DUMP=gpurun_out/infer_launches_new.json timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_new.json 2> gpurun_out/bench_new.err; echo "bench rc=$?"
