#!/bin/bash
# shared-memory stack bottom in the persistent extend kernel; C3 through the wavefront for comparison
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02_run21_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_run21_pytest.log
run() { python bench.py "${@:2}" --no-cpu-baseline --no-all-workloads 2>gpurun_out/r02_run21_$1.err | python -c "
import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{\"metric')][-1]); print('$1', round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],1), 'ms', d['mean_radiance'], d['gpu_launches'], 'e2e', round(d['e2e']['value'],1))"; }
C5="--workload C5 --spp 64 --steps 1 --warmup 1 --warmup-spp 2 --fused-e2e"
WRT_WF_SMSTACK=0 run sm0 $C5
WRT_WF_SMSTACK=4 run sm4 $C5
WRT_WF_SMSTACK=6 run sm6 $C5
WRT_WF_SMSTACK=8 run sm8 $C5
run c3_mega --workload C3 --steps 2 --warmup 1
run c3_wf --workload C3 --steps 2 --warmup 1 --engine wavefront
WRT_WIDE_TREE=1 run c3_wf_wide --workload C3 --steps 2 --warmup 1 --engine wavefront
