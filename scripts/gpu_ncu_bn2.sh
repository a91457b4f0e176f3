#!/bin/bash
mkdir -p gpurun_out
python scripts/profile_train.py 64 2 > gpurun_out/plain_train.log 2>&1 || { echo plain failed; tail -5 gpurun_out/plain_train.log; exit 1; }
ncu --set full --import-source on --clock-control none -k 'regex:^act_bn_bwd_kernel' -s 192 -c 1 -o gpurun_out/full_abb -f python scripts/profile_train.py 64 2 > gpurun_out/ncu_full_abb.log 2>&1
echo "abb exit=$?"
ncu --set full --import-source on --clock-control none -k 'regex:^bn_act_kernel' -s 107 -c 1 -o gpurun_out/full_bnact -f python scripts/profile_train.py 64 2 > gpurun_out/ncu_full_bnact.log 2>&1
echo "bnact exit=$?"
