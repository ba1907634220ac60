#!/bin/bash
# round 2, GPU call 6: more stage barriers, 32-warp blocks; then the bench at 50 Mb with the new defaults
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2f_pytest_gpu.log 2>&1; echo "pytest=$?"; tail -3 gpurun_out/r2f_pytest_gpu.log
cuobjdump -elf daccord_b200/_build/libdaccord_b200.so | grep '\$_ZN' > gpurun_out/r2f_symbols.txt
q() { local label=$1; shift; echo "$label: $(env "$@" timeout 300 python tools/ncu_target.py 10 40 3 2>&1 | grep kernel | tail -1 | cut -c1-100)"; }
q base X=1
for m in 15 63 111 127 175 239 255; do q mask$m DCU_SYNC_MASK=$m; done
q w32_g32 DCU_WPB=32 DCU_SYNC_GROUP=32
q w32_g16 DCU_WPB=32 DCU_SYNC_GROUP=16
q w32_g32_m255 DCU_WPB=32 DCU_SYNC_GROUP=32 DCU_SYNC_MASK=255
timeout 1200 python bench.py --steps 5 --warmup 3 --cpu-sample-s 4 2>gpurun_out/r2f_bench50.err > gpurun_out/r2f_bench50.json; python -c "
import json; l=json.load(open('gpurun_out/r2f_bench50.json')); print('bench50 value %.3f e2e %.3f cli %s' % (l['value']/1e6, l['e2e']['value']/1e6, (l.get('e2e_cli') or {}).get('value')))"
