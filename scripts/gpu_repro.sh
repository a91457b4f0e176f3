#!/bin/bash
for cfg in "8 3" "32 3" "64 3" "128 3"; do
  set -- $cfg
  timeout 120 python scripts/profile_fwd.py $1 $2 > gpurun_out/repro_$1.log 2>&1; echo "B=$1 n=$2 exit=$? $(tail -1 gpurun_out/repro_$1.log | cut -c1-150)"
done
