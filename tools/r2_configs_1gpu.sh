#!/bin/bash
# BASELINE configs on one GPU, scaled to the GPU-minute budget: config 4 (k sweep, 20 Mb per k instead of 200 Mb), the shallow tail (10x, 20x),
# config 5's pile shape on one GPU (200x, -d200 -D5000, 20 % tandem repeats, 3 Mb) and the single-GPU point of the strong-scaling run (config 3)
set -u
mkdir -p gpurun_out
run() { # name args...
  local name=$1; shift
  timeout 1200 python bench.py "$@" 2>gpurun_out/cfg_$name.err > gpurun_out/cfg_$name.json
  python -c "
import json,sys
l=json.load(open(sys.argv[1])); c=l.get('cpu_baseline') or {}
print(sys.argv[2], 'value %.3f M e2e %.3f M  cpu %.1f k (%s thr) identical %s  hard %d lost %d  acc %s  bytes/win %.0f frac %.2e' % (l['value']/1e6, l['e2e']['value']/1e6, c.get('value',0)/1e3, c.get('cores'), c.get('gpu_results_identical_on_sample'), l['hard_windows'], l['lost_windows'], (l.get('accuracy') or {}).get('erate'), l['roofline']['bytes_per_window'], l['roofline']['frac']))" gpurun_out/cfg_$name.json $name || tail -3 gpurun_out/cfg_$name.err
}
for k in 6 8 10 12 14; do run k$k --mb 20 --k $k --steps 3 --warmup 3 --cpu-sample-s 3 --cli 0 --truth-reads 200; done
run cov10 --mb 10 --coverage 10 --steps 3 --warmup 3 --cpu-sample-s 3 --cli 0 --truth-reads 200
run cov20 --mb 10 --coverage 20 --steps 3 --warmup 3 --cpu-sample-s 3 --cli 0 --truth-reads 200
run deep200 --mb 3 --coverage 200 --depth-cap 200 --maxinput 5000 --repeat-frac 0.2 --steps 2 --warmup 3 --cpu-sample-s 3 --cli 0 --truth-reads 100
