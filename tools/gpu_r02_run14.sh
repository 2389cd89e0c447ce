#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q > gpurun_out/r02_run14_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_run14_pytest.log
run() { python bench.py "${@:2}" --no-cpu-baseline --no-all-workloads 2>gpurun_out/r02_run14_$1.err | python -c "
import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{\"metric')][-1]); print('$1', round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],1), 'ms', d['mean_radiance'], d['gpu_launches'])"; }
run c5_16 --workload C5 --spp 16 --steps 1 --warmup 1 --warmup-spp 2 --fused-e2e
run c5_64 --workload C5 --spp 64 --steps 1 --warmup 1 --warmup-spp 2 --fused-e2e
run c5_mk16 --workload C5 --engine megakernel --spp 16 --steps 1 --warmup 1 --warmup-spp 2 --fused-e2e
