#!/bin/bash
# 256-bit loads in the per-lane traversals: parity, then C5 (wavefront) and C3 (per-lane megakernel)
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q > gpurun_out/r02_run19_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_run19_pytest.log
run() { python bench.py "${@:2}" --no-cpu-baseline --no-all-workloads 2>gpurun_out/r02_run19_$1.err | python -c "
import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{\"metric')][-1]); print('$1', round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],1), 'ms', d['mean_radiance'], d['gpu_launches'], 'e2e', round(d['e2e']['value'],1))"; }
C5="--workload C5 --spp 64 --steps 1 --warmup 1 --warmup-spp 2 --fused-e2e"
WRT_WF_SORT=0 run c5_nosort $C5
WRT_WF_SORT=1 run c5_sort $C5
run c3 --workload C3 --steps 2 --warmup 1
run c2_256 --workload C2 --spp 256 --steps 2 --warmup 1
