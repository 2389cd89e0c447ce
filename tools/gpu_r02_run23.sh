#!/bin/bash
# compact 8-byte stack entries in the persistent extend kernel: parity (frames vs megakernel, config-size hits), then C5 on / off
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q > gpurun_out/r02_run23_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_run23_pytest.log
run() { python bench.py "${@:2}" --no-cpu-baseline --no-all-workloads 2>gpurun_out/r02_run23_$1.err | python -c "
import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{\"metric')][-1]); print('$1', round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],1), 'ms', d['mean_radiance'], d['gpu_launches'], 'e2e', round(d['e2e']['value'],1))"; }
C5="--workload C5 --spp 64 --steps 1 --warmup 1 --warmup-spp 2 --fused-e2e"
WRT_COMPACT_STACK=0 run full16 $C5
WRT_COMPACT_STACK=1 run compact8 $C5
WRT_COMPACT_STACK=1 WRT_WF_BLOCKS=6 run compact8_b6 $C5
