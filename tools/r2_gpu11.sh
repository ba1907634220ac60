#!/bin/bash
# round 2, GPU call 11: next k-mer's probe in flight, scan prefetch, shuffle rank sort; VS table in global memory / carveout A/B; bench at 50 Mb with the CLI changes
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2k_pytest_gpu.log 2>&1; echo "pytest=$?"; tail -3 gpurun_out/r2k_pytest_gpu.log
ab() { local name=$1 mb=$2 cov=$3; shift 3; env "$@" timeout 600 python bench.py --mb $mb --coverage $cov --steps 3 --warmup 3 --cpu-sample-s 0 --cli 0 --truth-reads 0 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read()); print('$name value %.3f e2e %.3f hard %d launches %d' % (l['value']/1e6, l['e2e']['value']/1e6, l['hard_windows'], l['gpu_launches']))"; }
ab base40 10 40 X=1
ab vsglobal40 10 40 DCU_VS_GLOBAL=1
ab carve35 10 40 DCU_CARVEOUT=35
ab base20 10 20 X=1
ab base10 5 10 X=1
timeout 1200 python bench.py --steps 5 --warmup 3 --cpu-sample-s 4 2>gpurun_out/r2k_bench50.err > gpurun_out/r2k_bench50.json; python -c "
import json; l=json.load(open('gpurun_out/r2k_bench50.json')); print('bench50 value %.3f e2e %.3f hard %d launches %d cli %s' % (l['value']/1e6, l['e2e']['value']/1e6, l['hard_windows'], l['gpu_launches'], l.get('e2e_cli')))"
