#!/bin/bash
# C5: record numbering (breadth first vs depth first, host build), prefetch of deferred children, reordering on top
set -u
mkdir -p gpurun_out
run() { python bench.py "${@:2}" --no-cpu-baseline --no-all-workloads 2>gpurun_out/r02_run18_$1.err | python -c "
import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{\"metric')][-1]); print('$1', round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],1), 'ms', d['mean_radiance'], d['gpu_launches'], 'e2e', round(d['e2e']['value'],1))"; }
C5="--workload C5 --spp 64 --steps 1 --warmup 1 --warmup-spp 2 --fused-e2e"
WRT_WF_SORT=0 WRT_DEVICE_BUILD=0 run bfs_host $C5
WRT_WF_SORT=0 WRT_DEVICE_BUILD=0 WRT_NODE4_ORDER=dfs run dfs_host $C5
WRT_WF_SORT=0 WRT_TRAV_PREFETCH=1 run pf1 $C5
WRT_WF_SORT=0 WRT_TRAV_PREFETCH=2 run pf2 $C5
WRT_WF_SORT=0 WRT_TRAV_PREFETCH=3 run pf3 $C5
WRT_WF_SORT=1 WRT_TRAV_PREFETCH=2 run sort_pf2 $C5
