#!/bin/bash
# round 2, GPU call 17: reverse slots by (link, position) pairs in two phases (DCU_SPFLAT) against one link at a time; tests, bench, ncu capture of the head
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2q_pytest_gpu.log 2>&1; echo "pytest=$?"; tail -2 gpurun_out/r2q_pytest_gpu.log
ab() { local name=$1; shift; env "$@" timeout 300 python bench.py --mb 10 --steps 3 --warmup 3 --cpu-sample-s 0 --cli 0 --truth-reads 0 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read()); print('$name value %.3f e2e %.3f' % (l['value']/1e6, l['e2e']['value']/1e6))"; }
ab flat1 DCU_SPFLAT=1
ab flat0 DCU_SPFLAT=0
ab flat1b DCU_SPFLAT=1
ab flat0b DCU_SPFLAT=0
timeout 900 python bench.py --steps 5 --warmup 3 --cpu-sample-s 4 2>gpurun_out/r2q_bench.err > gpurun_out/r2q_bench.json; python -c "
import json
l=json.load(open('gpurun_out/r2q_bench.json'))
print('bench50 value %.3f M e2e %.3f M two %.3f cli %s identical %s' % (l['value']/1e6, l['e2e']['value']/1e6, l['e2e_two_in_flight']['value']/1e6, l['e2e_cli']['value'], l['cpu_baseline']['gpu_results_identical_on_sample']))"
cuobjdump -elf daccord_b200/_build/libdaccord_b200.so | grep '\$_ZN' > gpurun_out/r2q_symbols.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dcu_window -s 1 -c 1 -f -o gpurun_out/r2q_head_full python tools/ncu_target.py 2 40 2 > gpurun_out/r2q_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r2q_head_full.ncu-rep --page raw --csv > gpurun_out/r2q_head_raw.csv 2>/dev/null
ncu -i gpurun_out/r2q_head_full.ncu-rep --page source --csv > gpurun_out/r2q_head_source.csv 2>/dev/null
rm -f gpurun_out/r2q_head_full.ncu-rep
