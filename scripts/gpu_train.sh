#!/bin/bash
# training-path iteration: op tests + whole-model training parity
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 python -m pytest "$@" -q --timeout 600 -p no:cacheprovider > gpurun_out/test_$name.log 2>&1; echo "$name exit=$? $(tail -1 gpurun_out/test_$name.log)"; }
run train_ops tests/test_gpu_train_ops.py
run train_model tests/test_gpu_train_model.py -s
grep -E "FAILED|Error|error|assert" gpurun_out/test_train_ops.log | head -30
grep -E "FAILED|Error|error|assert|worst|cosine" gpurun_out/test_train_model.log | head -30
