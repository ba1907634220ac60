#!/bin/bash
# round 2, GPU call 14: host-side laps of dcu_pile / dcu_vote on the 50 Mb workload (DCU_TIMING)
export DCU_TIMING=1
timeout 900 python bench.py --steps 2 --warmup 3 --cpu-sample-s 0 --cli 0 --truth-reads 0 2>&1 >/dev/null | grep timing | tail -12
