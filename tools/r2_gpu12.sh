#!/bin/bash
# round 2, GPU call 12: heavy worker blocks (windows with many pairs leave the lock step and are redone by free-running blocks of the same launch)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2l_pytest_gpu.log 2>&1; echo "pytest=$?"; tail -3 gpurun_out/r2l_pytest_gpu.log
ab() { local name=$1 mb=$2 cov=$3; shift 3; env "$@" timeout 600 python bench.py --mb $mb --coverage $cov --steps 3 --warmup 3 --cpu-sample-s 0 --cli 0 --truth-reads 0 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read()); print('$name value %.3f e2e %.3f hard %d launches %d' % (l['value']/1e6, l['e2e']['value']/1e6, l['hard_windows'], l['gpu_launches']))"; }
ab off40 10 40 DCU_HEAVY_BLOCKS=0
ab hb8_p32 10 40 X=1
ab hb8_p16 10 40 DCU_MAXPAIRS=16
ab hb8_p64 10 40 DCU_MAXPAIRS=64
ab hb16_p32 10 40 DCU_HEAVY_BLOCKS=16
ab hb4_p32 10 40 DCU_HEAVY_BLOCKS=4
ab off20 10 20 DCU_HEAVY_BLOCKS=0
ab hb8_20 10 20 X=1
ab off10 5 10 DCU_HEAVY_BLOCKS=0
ab hb8_10 5 10 X=1
ab hb8_50mb 50 40 X=1
