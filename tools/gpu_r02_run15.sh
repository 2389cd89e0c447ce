#!/bin/bash
set -u
mkdir -p gpurun_out
run() { python bench.py "${@:2}" --no-cpu-baseline --no-all-workloads 2>gpurun_out/r02_run15_$1.err | python -c "
import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{\"metric')][-1]); print('$1', round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],1), 'ms', d['mean_radiance'], d['gpu_launches'])"; }
for pool in 4 16 32 64; do
WRT_WF_POOL=$pool run c5_64_pool$pool --workload C5 --spp 64 --steps 1 --warmup 1 --warmup-spp 2 --fused-e2e
done
