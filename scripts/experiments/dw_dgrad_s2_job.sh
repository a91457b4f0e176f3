#!/bin/bash
# Same-box A/B of the stride-2 depthwise dgrad (previous build vs this build), then the GPU suite and the bench line on this build.
# scripts/experiments/_ab/libdfvit_prev.so = the build to compare against (git stash / checkout the older sources, make, copy the .so there;
# *.so is git-ignored but travels to the GPU box).
mkdir -p gpurun_out
cp deepfake_vit_b200/libdfvit.so /tmp/libdfvit_new.so
cp scripts/experiments/_ab/libdfvit_prev.so deepfake_vit_b200/libdfvit.so
timeout 120 python scripts/experiments/dw_dgrad_s2_ab.py prev > gpurun_out/dgrad_ab_prev.log 2>&1; echo "ab prev rc=$?"
cp /tmp/libdfvit_new.so deepfake_vit_b200/libdfvit.so
timeout 120 python scripts/experiments/dw_dgrad_s2_ab.py new prev > gpurun_out/dgrad_ab_new.log 2>&1; echo "ab new rc=$?"
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_final.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/gpu_tests_final.log
DFV_BENCH_DUMP=gpurun_out/launches_final.json timeout 200 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
