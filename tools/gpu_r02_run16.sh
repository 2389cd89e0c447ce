#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q > gpurun_out/r02_run16_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_run16_pytest.log
run() { python bench.py "${@:2}" --no-cpu-baseline --no-all-workloads 2>gpurun_out/r02_run16_$1.err | python -c "
import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{\"metric')][-1]); print('$1', round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],1), 'ms', d['mean_radiance'], d['gpu_launches'])"; }
for cfg in "1 32" "2 32" "3 32" "2 16" "4 64"; do
set -- $cfg
WRT_WF_PIPELINES=$1 WRT_WF_POOL=$2 run c5_64_p$1_pool$2 --workload C5 --spp 64 --steps 1 --warmup 1 --warmup-spp 2 --fused-e2e
done
