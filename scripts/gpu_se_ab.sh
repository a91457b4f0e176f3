#!/bin/bash
# in-pipeline cost of the squeeze-excite gate: step time with it and with its launches skipped (DFV_DEBUG_FLAGS=2)
run() { python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(round(j['ms_per_step'],3))"; }
for i in 1 2; do
  echo -n "full: "; run
  echo -n "skip launch: "; DFV_DEBUG_FLAGS=2 run
done
