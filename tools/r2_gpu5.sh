#!/bin/bash
# round 2, GPU call 5: pipelined table probes; sweep of stage barriers, group sizes, warps per block and first-pass capacities on the 10 Mb pile (HBM build)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2e_pytest_gpu.log 2>&1; echo "pytest=$?"; tail -3 gpurun_out/r2e_pytest_gpu.log
export DCU_NO_SMEM=1
q() { # label env...
  local label=$1; shift
  echo "$label: $(env "$@" timeout 300 python tools/ncu_target.py 10 40 3 2>&1 | grep kernel | tail -1 | cut -c1-120)"
}
q base X=1
for m in 0 8 4 12 6 14 7 31 47 79; do q mask$m DCU_SYNC_MASK=$m; done
q g8 DCU_SYNC_GROUP=8
q wpb12 DCU_WPB=12
q wpb12_g6 DCU_WPB=12 DCU_SYNC_GROUP=4
q small DCU_T0_SMALL=1
q small_wpb12 DCU_T0_SMALL=1 DCU_WPB=12
q bps1 DCU_BLOCKS_PER_SM=1
q cov20 X=1
echo "cov20 base: $(timeout 300 python tools/ncu_target.py 10 20 3 2>&1 | grep kernel | tail -1 | cut -c1-120)"
echo "cov20 g16: $(DCU_SYNC_GROUP=16 timeout 300 python tools/ncu_target.py 10 20 3 2>&1 | grep kernel | tail -1 | cut -c1-120)"
echo "cov10 base: $(timeout 300 python tools/ncu_target.py 5 10 3 2>&1 | grep kernel | tail -1 | cut -c1-120)"
echo "cov10 g8: $(DCU_SYNC_GROUP=8 timeout 300 python tools/ncu_target.py 5 10 3 2>&1 | grep kernel | tail -1 | cut -c1-120)"
echo "cov10 g16: $(DCU_SYNC_GROUP=16 timeout 300 python tools/ncu_target.py 5 10 3 2>&1 | grep kernel | tail -1 | cut -c1-120)"
