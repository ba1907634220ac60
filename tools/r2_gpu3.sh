#!/bin/bash
# round 2, GPU call 3: lean sp_view (per-lane instance loops, shared-memory table, wf/wl slots) + rolled k-mer passes: tests, bench A/B of the two builds, ncu of both
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2c_pytest_gpu.log 2>&1; echo "pytest=$?"; tail -5 gpurun_out/r2c_pytest_gpu.log
run() { # name mb cov env...
  local name=$1 mb=$2 cov=$3; shift 3
  env "$@" timeout 600 python bench.py --mb $mb --coverage $cov --steps 3 --warmup 3 --cpu-sample-s 0 --truth-reads 100 2>gpurun_out/r2c_$name.err > gpurun_out/r2c_$name.json
  python -c "import json,sys; l=json.load(open(sys.argv[1])); print(sys.argv[2],'value %.3f e2e %.3f desc %.3f second %d hard %d lost %d smem %s acc %s' % (l['value']/1e6, l['e2e']['value']/1e6, l['e2e_descriptors']['value']/1e6, l['second_pass_windows'], l['hard_windows'], l['lost_windows'], l['smem_pass'], l['accuracy']))" gpurun_out/r2c_$name.json $name || tail -3 gpurun_out/r2c_$name.err
}
run smem40 10 40 X=1
run hbm40 10 40 DCU_NO_SMEM=1
run smem40_g4 10 40 DCU_SYNC_GROUP=4
run smem40_g1 10 40 DCU_SYNC_GROUP=1
run hbm40_g8 10 40 DCU_NO_SMEM=1 DCU_SYNC_GROUP=8
run smem20 10 20 X=1
run hbm20 10 20 DCU_NO_SMEM=1
run smem10 5 10 X=1
run hbm10 5 10 DCU_NO_SMEM=1
for b in smem hbm; do
  if [ $b = hbm ]; then export DCU_NO_SMEM=1; K=dcu_window; else unset DCU_NO_SMEM; K=dcus_window; fi
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/r2c_${b}_full python tools/ncu_target.py 2 40 2 > gpurun_out/r2c_ncu_$b.log 2>&1; echo "ncu $b rc=$?"; tail -2 gpurun_out/r2c_ncu_$b.log
  ncu -i gpurun_out/r2c_${b}_full.ncu-rep --page raw --csv > gpurun_out/r2c_${b}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r2c_${b}_full.ncu-rep --page source --csv > gpurun_out/r2c_${b}_source.csv 2>/dev/null
  rm -f gpurun_out/r2c_${b}_full.ncu-rep
done
unset DCU_NO_SMEM
