#!/bin/bash
# round 2, GPU call 7: A/B on the bench harness after taking the 4-wide probes out again (stage barriers, 32-warp blocks), then the BASELINE configs on one GPU
set -u
mkdir -p gpurun_out
cuobjdump -elf daccord_b200/_build/libdaccord_b200.so | grep '\$_ZN' > gpurun_out/r2g_symbols.txt
ab() { local name=$1; shift; env "$@" timeout 600 python bench.py --mb 10 --steps 3 --warmup 3 --cpu-sample-s 0 --cli 0 --truth-reads 0 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read()); print('$name value %.3f e2e %.3f' % (l['value']/1e6, l['e2e']['value']/1e6))"; }
ab mask255 X=1
ab mask47 DCU_SYNC_MASK=47
ab mask15 DCU_SYNC_MASK=15
ab w32_m47 DCU_WPB=32 DCU_SYNC_GROUP=32 DCU_SYNC_MASK=47
ab w32_m255 DCU_WPB=32 DCU_SYNC_GROUP=32
ab smem DCU_SMEM=1
bash tools/r2_configs_1gpu.sh
