#!/bin/bash
# First GPU call of the next round (about 15 GPU-minutes): everything that could not be measured when round 1 ran out of budget.
#   gpurun --timeout 1500 -- 'bash tools/round2_first_run.sh'
set -u
mkdir -p gpurun_out
# 1. the whole GPU suite: the per-window piling kernels, the 32-bit unitig slot offsets, the position-slot cache and the binary-search
#    pairing range went in after the last GPU run (emulation parity only: single lane, 32 lanes under adversarial schedules, TSan)
timeout 500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest=$?"; tail -3 gpurun_out/r2_pytest_gpu.log
# 1b. memcheck of a small mixed batch through every kernel: the changes made without a GPU include 8-byte slot / pattern loads and
#     new workspace fields (alignment and bounds are what the host emulations cannot see)
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/r2_memcheck.log 2>&1; echo "memcheck=$?"; grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/r2_memcheck.log; tail -2 gpurun_out/r2_memcheck.log
line() { python -c "import json,sys; l=json.load(open(sys.argv[1])); print(sys.argv[2], 'value %.3f e2e %.3f from_overlaps %.3f to_fasta %.3f hard %d' % (l['value']/1e6, l['e2e']['value']/1e6, l['e2e_from_overlaps']['value']/1e6, l['e2e_overlaps_to_fasta']['value']/1e6, l['hard_windows']))" "$1" "$2"; }
# 2. A/B of the position-slot cache (profiles/r01_summary.md) on the bench workload at 40x and 20x
for pc in 1 0; do
  for cov in 40 20; do
    DCU_POSCACHE=$pc python bench.py --mb 20 --coverage $cov --steps 3 --warmup 2 --cpu-sample-s 0 2>/dev/null > gpurun_out/r2_poscache_${pc}_cov${cov}.json
    line gpurun_out/r2_poscache_${pc}_cov${cov}.json "poscache=$pc coverage=$cov"
  done
done
# 3. A/B of the deferral switch
DCU_DEFER_FF=1 python bench.py --mb 20 --steps 3 --warmup 2 --cpu-sample-s 0 2>/dev/null > gpurun_out/r2_defer_1.json; line gpurun_out/r2_defer_1.json "defer=1"
# 4. resident warps per SM against the L2 (4 736 workspaces of ~64-100 KB do not fit 126 MB): 1 block of 16 warps per SM
DCU_BLOCKS_PER_SM=1 python bench.py --mb 20 --steps 3 --warmup 2 --cpu-sample-s 0 2>/dev/null > gpurun_out/r2_bps_1.json; line gpurun_out/r2_bps_1.json "blocks_per_sm=1"
# 4b. first-pass capacities near the workload's p99.9 (slab 399 KB instead of 950 KB per warp: fewer live 2 MB pages per SM, more second-pass windows)
DCU_T0_SMALL=1 python bench.py --mb 20 --steps 3 --warmup 2 --cpu-sample-s 0 2>/dev/null > gpurun_out/r2_t0small.json; line gpurun_out/r2_t0small.json "t0_small=1"
# 4c. group size at 20x again: the cache removes most of the heavy tail that made small groups win on shallow piles (profiles/r01_summary.md, sync-group sweep)
for g in 16 8 1; do DCU_SYNC_GROUP=$g python bench.py --mb 5 --coverage 10 --steps 2 --warmup 2 --cpu-sample-s 0 2>/dev/null > gpurun_out/r2_group_${g}_cov10.json; line gpurun_out/r2_group_${g}_cov10.json "sync_group=$g coverage=10"; done
for g in 16 8 1; do DCU_SYNC_GROUP=$g python bench.py --mb 10 --coverage 20 --steps 2 --warmup 2 --cpu-sample-s 0 2>/dev/null > gpurun_out/r2_group_${g}_cov20.json; line gpurun_out/r2_group_${g}_cov20.json "sync_group=$g coverage=20"; done
# 5. hard configurations the cache was written for (tools/kernel_bench.py: synthetic windows, device-timed)
for pc in 1 0; do DCU_POSCACHE=$pc python tools/kernel_bench.py 10 20 2>&1 | tail -1 | sed "s/^/poscache=$pc depth10 /"; done
# 6. launch list of the bench command with the new piling kernels
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_5mb.csv python bench.py --mb 5 --steps 2 --warmup 1 --cpu-sample-s 0 > /dev/null 2>&1; echo "launch list rc=$?"
