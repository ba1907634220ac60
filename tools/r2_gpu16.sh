#!/bin/bash
# round 2, GPU call 16: GPU piling for any -w / -a; GPU tests and the default bench line once more
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2p_pytest_gpu.log 2>&1; echo "pytest=$?"; tail -3 gpurun_out/r2p_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2p_smoke.log 2>&1; echo "smoke=$?"; tail -1 gpurun_out/r2p_smoke.log
timeout 1500 python bench.py --steps 5 --warmup 3 --cpu-sample-s 4 2>gpurun_out/r2p_bench.err > gpurun_out/r2p_bench.json; echo "bench=$?"
python -c "
import json
l=json.load(open('gpurun_out/r2p_bench.json'))
print('value %.3f M e2e %.3f M (%s) two %.3f cli %s' % (l['value']/1e6, l['e2e']['value']/1e6, l['e2e']['rank0_ms_per_step'], l['e2e_two_in_flight']['value']/1e6, l['e2e_cli']['value']))"
timeout 300 python bench.py --mb 10 --a 7 --steps 2 --warmup 3 --cpu-sample-s 3 --cli 1 --truth-reads 100 2>gpurun_out/r2p_a7.err > gpurun_out/r2p_a7.json; python -c "
import json
l=json.load(open('gpurun_out/r2p_a7.json'))
print('a7: value %.3f M e2e %.3f M identical %s fasta %s cli %s acc %s' % (l['value']/1e6, l['e2e']['value']/1e6, l['cpu_baseline']['gpu_results_identical_on_sample'], l['e2e']['fasta_identical_to_host_vote'], {k: l['e2e_cli'][k] for k in ('value','fasta_identical_to_library_path','fasta_identical_to_oracle_file_driver') if k in l['e2e_cli']}, l['accuracy']['erate']))"
