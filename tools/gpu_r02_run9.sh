#!/bin/bash
set -u
mkdir -p gpurun_out
CMD="python bench.py --workload C5 --spp 16 --engine wavefront --steps 1 --warmup 0 --fused-e2e --no-cpu-baseline --no-all-workloads"
$CMD > gpurun_out/r02_run9_plain.json 2> gpurun_out/r02_run9_plain.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_run9_launches.csv $CMD > gpurun_out/r02_run9_ncu.log 2>&1; echo "ncu rc=$?"
