#!/bin/bash
# Same-box A/B of two builds of libdfvit.so: bench line + per-launch dump under each.  Usage: bash scripts/ab_lib.sh <other.so> <tag>
mkdir -p gpurun_out
cp deepfake_vit_b200/libdfvit.so /tmp/libdfvit_head.so
for round in 1 2; do
  cp "$1" deepfake_vit_b200/libdfvit.so
  DFV_BENCH_DUMP=gpurun_out/launches_$2_$round.json timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$2_$round.json 2> gpurun_out/bench_$2_$round.err; echo "bench $2 rc=$?"
  cp /tmp/libdfvit_head.so deepfake_vit_b200/libdfvit.so
  DFV_BENCH_DUMP=gpurun_out/launches_head_$round.json timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_head_$round.json 2> gpurun_out/bench_head_$round.err; echo "bench head rc=$?"
done
