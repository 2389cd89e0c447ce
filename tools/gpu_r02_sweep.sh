#!/bin/bash
set -u
mkdir -p gpurun_out
for v in A B C D E F; do
  lib=""; [ $v != A ] && lib="$PWD/scratch/variants/libwrt_$v.so"
  WRT_LIB=$lib python bench.py --workload C5 --spp 32 --steps 1 --warmup 1 --warmup-spp 2 --fused-e2e --no-cpu-baseline --no-all-workloads > gpurun_out/r02_sweep_$v.json 2> gpurun_out/r02_sweep_$v.err
  python - <<P
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02_sweep_$v.json').read().splitlines() if l.startswith('{"metric')][-1]); print('$v', round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],1), 'ms', d['mean_radiance'])
except Exception as e: print('$v ERR', e)
P
done
