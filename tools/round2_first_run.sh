#!/bin/bash
# First GPU call of the next round (about 6 GPU-minutes): everything that could not be measured when round 1 ran out of budget.
#   gpurun --timeout 900 -- 'bash tools/round2_first_run.sh'
set -u
mkdir -p gpurun_out
# 1. the whole GPU suite: three tests were not reached after the per-window piling kernels went in, and the window kernel's unitig position
#    offsets were widened to 32 bit afterwards (emulation parity only)
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest=$?"; tail -3 gpurun_out/r2_pytest_gpu.log
# 2. A/B of the deferral switch on the bench workload
for d in 0 1; do
  if [ $d = 1 ]; then export DCU_DEFER_FF=1; else unset DCU_DEFER_FF; fi
  python bench.py --mb 20 --steps 3 --warmup 2 --cpu-sample-s 0 2>/dev/null > gpurun_out/r2_defer_$d.json
  python -c "import json; l=json.load(open('gpurun_out/r2_defer_$d.json')); print('defer=$d value %.3f e2e %.3f from_overlaps %.3f to_fasta %.3f hard %d' % (l['value']/1e6, l['e2e']['value']/1e6, l['e2e_from_overlaps']['value']/1e6, l['e2e_overlaps_to_fasta']['value']/1e6, l['hard_windows']))"
done
unset DCU_DEFER_FF
# 3. launch list of the bench command with the new piling kernels
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_5mb.csv python bench.py --mb 5 --steps 2 --warmup 1 --cpu-sample-s 0 > /dev/null 2>&1; echo "launch list rc=$?"
