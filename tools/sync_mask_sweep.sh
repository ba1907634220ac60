#!/bin/bash
# sweep of the stage barriers of dcu_window_kernel (DCU_SYNC_MASK bits: 1 hash|nodes, 32 nodes|edges, 2 edges|trav, 16 trav|pos, 4 pos|rpath,
# 64 rpath|search, 128 search|score, 8 score|final) on the bench workload; prints Mwin/s per setting
MB=${MB:-20}
for m in 15 47 79 143 239 255; do
  v=$(DCU_SYNC_MASK=$m python bench.py --mb $MB --steps 3 --warmup 2 --cpu-sample-s 0 2>/dev/null | python -c "import sys,json; l=json.loads(sys.stdin.read()); print('%.3f e2e %.3f'%(l['value']/1e6, l['e2e']['value']/1e6))")
  echo "bench${MB}mb mask=$m value $v"
done
for d in 40 20; do for m in 15 239; do
  echo "synthetic depth=$d mask=$m: $(DCU_SYNC_MASK=$m python tools/kernel_bench.py $d 100 | grep kernel | tail -1)"
done; done
