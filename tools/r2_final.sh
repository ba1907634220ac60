#!/bin/bash
# round 2, final GPU call at the head commit: GPU tests, the default bench line and the reference arm, the ncu capture the roofline is read from (with the
# symbol table of the profiled library), the launch list of the bench command
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2z_pytest_gpu.log 2>&1; echo "pytest=$?"; tail -3 gpurun_out/r2z_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; echo "smoke=$?"; tail -2 gpurun_out/r2z_smoke.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/r2z_memcheck.log 2>&1; echo "memcheck=$?"; grep -E "ERROR SUMMARY|pile" gpurun_out/r2z_memcheck.log | tail -2
cuobjdump -elf daccord_b200/_build/libdaccord_b200.so | grep '\$_ZN' > gpurun_out/r2z_symbols.txt
timeout 1500 python bench.py --steps 10 --warmup 5 2>gpurun_out/r2z_bench.err > gpurun_out/r2z_bench.json; echo "bench=$?"
timeout 600 python bench.py --impl reference --steps 10 --warmup 5 2>gpurun_out/r2z_ref.err > gpurun_out/r2z_ref.json; echo "reference=$?"
python -c "
import json
l=json.load(open('gpurun_out/r2z_bench.json')); r=json.load(open('gpurun_out/r2z_ref.json'))
print('value %.3f M e2e %.3f M cli %s reference %.1f k (%s threads) cpu_baseline %s' % (l['value']/1e6, l['e2e']['value']/1e6, (l.get('e2e_cli') or {}).get('value'), r['value']/1e3, r['config']['threads'], l.get('cpu_baseline')))"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dcu_window -s 1 -c 1 -f -o gpurun_out/r2z_head_full python tools/ncu_target.py 2 40 2 > gpurun_out/r2z_ncu.log 2>&1; echo "ncu rc=$?"; grep coverage gpurun_out/r2z_ncu.log | head -2
ncu -i gpurun_out/r2z_head_full.ncu-rep --page raw --csv > gpurun_out/r2z_head_raw.csv 2>/dev/null
ncu -i gpurun_out/r2z_head_full.ncu-rep --page source --csv > gpurun_out/r2z_head_source.csv 2>/dev/null
rm -f gpurun_out/r2z_head_full.ncu-rep
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2z_launches_5mb.csv python bench.py --mb 5 --steps 2 --warmup 1 --cpu-sample-s 0 --cli 0 > /dev/null 2>&1; echo "launch list rc=$?"
