#!/bin/bash
# round 2, GPU call 2: the two-build kernel (shared-memory first pass) -- tests, bench A/B against the HBM-only path, full ncu capture, launch list
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2b_pytest_gpu.log 2>&1; echo "pytest=$?"; tail -5 gpurun_out/r2b_pytest_gpu.log
run() { # name mb cov env...
  local name=$1 mb=$2 cov=$3; shift 3
  env "$@" timeout 500 python bench.py --mb $mb --coverage $cov --steps 3 --warmup 3 --cpu-sample-s 0 2>gpurun_out/r2b_$name.err > gpurun_out/r2b_$name.json
  python -c "import json,sys; l=json.load(open(sys.argv[1])); print(sys.argv[2],'value %.3f e2e %.3f to_fasta %.3f second %d hard %d smem %s' % (l['value']/1e6, l['e2e']['value']/1e6, l['e2e_overlaps_to_fasta']['value']/1e6, l['second_pass_windows'], l['hard_windows'], l['smem_pass']))" gpurun_out/r2b_$name.json $name || tail -3 gpurun_out/r2b_$name.err
}
run smem40 10 40 X=1
run hbm40 10 40 DCU_NO_SMEM=1
run smem40_nostage 10 40 DCU_STAGE=0
run smem20 10 20 X=1
run smem10 5 10 X=1
run hbm10 5 10 DCU_NO_SMEM=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dcus_window -s 1 -c 1 -f -o gpurun_out/r2b_smem_full python tools/ncu_target.py 2 40 2 > gpurun_out/r2b_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/r2b_ncu_full.log
ncu -i gpurun_out/r2b_smem_full.ncu-rep --page raw --csv > gpurun_out/r2b_smem_raw.csv 2>/dev/null
ncu -i gpurun_out/r2b_smem_full.ncu-rep --page source --csv > gpurun_out/r2b_smem_source.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2b_launches_5mb.csv python bench.py --mb 5 --steps 2 --warmup 1 --cpu-sample-s 0 > /dev/null 2>&1; echo "launch list rc=$?"
