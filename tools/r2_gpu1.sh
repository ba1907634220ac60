#!/bin/bash
# round 2, first GPU call: head of round 1 re-measured (tests, bench at three coverages, full ncu capture on bench-workload windows, launch list)
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/r2a_pytest_gpu.log 2>&1; echo "pytest=$?"; tail -2 gpurun_out/r2a_pytest_gpu.log
for cfg in "10 40" "10 20" "5 10"; do set -- $cfg
  timeout 400 python bench.py --mb $1 --coverage $2 --steps 3 --warmup 3 --cpu-sample-s 0 2>gpurun_out/r2a_bench_cov$2.err > gpurun_out/r2a_bench_cov$2.json
  python -c "import json,sys; l=json.load(open(sys.argv[1])); print('cov',sys.argv[2],'value %.3f e2e %.3f to_fasta %.3f hard %d' % (l['value']/1e6, l['e2e']['value']/1e6, l['e2e_overlaps_to_fasta']['value']/1e6, l['hard_windows']))" gpurun_out/r2a_bench_cov$2.json $2
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dcu_window -s 1 -c 1 -f -o gpurun_out/r2a_head_full python tools/ncu_target.py 2 40 2 > gpurun_out/r2a_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/r2a_ncu_full.log
ncu -i gpurun_out/r2a_head_full.ncu-rep --page raw --csv > gpurun_out/r2a_head_raw.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2a_launches_5mb.csv python bench.py --mb 5 --steps 2 --warmup 1 --cpu-sample-s 0 > /dev/null 2>&1; echo "launch list rc=$?"
