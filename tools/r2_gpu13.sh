#!/bin/bash
# round 2, GPU call 13 (diagnostic build): when do the warps of a launch finish their last window?
set -u
export DCU_DEBUG_TIMES=1
for cfg in "2 40" "10 40" "10 20"; do set -- $cfg
  echo "== $1 Mb, ${2}x"; timeout 300 python tools/ncu_target.py $1 $2 3 2>&1 | grep -E "dbg|kernel" | cut -c1-330
done
