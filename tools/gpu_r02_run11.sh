#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_run11_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_run11_pytest.log
for v in on off; do
  p=1; [ $v = off ] && p=0
  WRT_L2_PERSIST=$p python bench.py --workload C5 --spp 32 --steps 1 --warmup 1 --warmup-spp 2 --fused-e2e --no-cpu-baseline --no-all-workloads > gpurun_out/r02_run11_l2$v.json 2> gpurun_out/r02_run11_l2$v.err
  python - <<P
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02_run11_l2$v.json').read().splitlines() if l.startswith('{"metric')][-1]); print('l2 persist $v', round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],1), 'ms', d['mean_radiance'])
except Exception as e: print('$v ERR', e)
P
done
python bench.py --workload C5 --engine megakernel --spp 32 --steps 1 --warmup 1 --warmup-spp 2 --fused-e2e --no-cpu-baseline --no-all-workloads 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{\"metric')][-1]); print('megakernel', round(d['value'],1), d['mean_radiance'])"
