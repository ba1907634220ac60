#!/bin/bash
# round 2, GPU call 4: slot records, register-resident placement, pipelined CLI; A/B of group sizes / first-pass capacities; ncu of the HBM build (first-pass launch)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2d_pytest_gpu.log 2>&1; echo "pytest=$?"; tail -5 gpurun_out/r2d_pytest_gpu.log
run() { # name mb cov extra-args env...
  local name=$1 mb=$2 cov=$3 extra=$4; shift 4
  env "$@" timeout 900 python bench.py --mb $mb --coverage $cov --steps 3 --warmup 3 --cpu-sample-s 0 --truth-reads 100 $extra 2>gpurun_out/r2d_$name.err > gpurun_out/r2d_$name.json
  python -c "import json,sys; l=json.load(open(sys.argv[1])); print(sys.argv[2],'value %.3f e2e %.3f desc %.3f second %d hard %d lost %d cli %s' % (l['value']/1e6, l['e2e']['value']/1e6, l['e2e_descriptors']['value']/1e6, l['second_pass_windows'], l['hard_windows'], l['lost_windows'], {k: v for k, v in (l.get('e2e_cli') or {}).items() if k not in ('what',)}))" gpurun_out/r2d_$name.json $name || tail -3 gpurun_out/r2d_$name.err
}
run hbm40 10 40 "--cli 0" DCU_NO_SMEM=1
run hbm40_small 10 40 "--cli 0" DCU_NO_SMEM=1 DCU_T0_SMALL=1
run smem40_g4 10 40 "--cli 0" DCU_SYNC_GROUP=4
run smem40_g2 10 40 "--cli 0" DCU_SYNC_GROUP=2
run smem40_g6 10 40 "--cli 0" DCU_SYNC_GROUP=8
run hbm50 50 40 "--cpu-sample-s 4" DCU_NO_SMEM=1
export DCU_NO_SMEM=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dcu_window -s 2 -c 1 -f -o gpurun_out/r2d_hbm_full python tools/ncu_target.py 2 40 2 > gpurun_out/r2d_ncu_hbm.log 2>&1; echo "ncu hbm rc=$?"; tail -2 gpurun_out/r2d_ncu_hbm.log
ncu -i gpurun_out/r2d_hbm_full.ncu-rep --page raw --csv > gpurun_out/r2d_hbm_raw.csv 2>/dev/null
ncu -i gpurun_out/r2d_hbm_full.ncu-rep --page source --csv > gpurun_out/r2d_hbm_source.csv 2>/dev/null
rm -f gpurun_out/r2d_hbm_full.ncu-rep
