#!/bin/bash
# Same-box A/B of the training-path attention kernels (previous build vs this build), then the GPU suite and the bench line.
# scripts/experiments/_ab/libdfvit_prev.so = the build to compare against (git stash / checkout the older sources, make, copy the .so there;
# *.so is git-ignored but travels to the GPU box).
mkdir -p gpurun_out
cp deepfake_vit_b200/libdfvit.so /tmp/libdfvit_new.so
cp scripts/experiments/_ab/libdfvit_prev.so deepfake_vit_b200/libdfvit.so
timeout 100 python scripts/experiments/attn_train_ab.py prev > gpurun_out/attn_ab_prev.log 2>&1; echo "ab prev rc=$?"
cp /tmp/libdfvit_new.so deepfake_vit_b200/libdfvit.so
timeout 100 python scripts/experiments/attn_train_ab.py new prev > gpurun_out/attn_ab_new.log 2>&1; echo "ab new rc=$?"
timeout 200 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_attn.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/gpu_tests_attn.log
DFV_BENCH_DUMP=gpurun_out/launches_attn.json timeout 150 python bench.py > gpurun_out/bench_attn.json 2> gpurun_out/bench_attn.err; echo "bench rc=$?"
