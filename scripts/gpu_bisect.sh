#!/bin/bash
for f in 0 7 8 16 24 1 2; do
  r=$(DFV_DEBUG_FLAGS=$f timeout 300 python scripts/profile_fwd.py 256 ${1:-60} 2>&1 | grep -E "FAILED|^ok" | cut -c1-160)
  echo "flags=$f: $r"
done
