#!/bin/bash
set -u
mkdir -p gpurun_out
CMD="python bench.py --workload C5 --res 960x540 --spp 16 --engine wavefront --steps 1 --warmup 0 --no-cpu-baseline --no-all-workloads"
$CMD > gpurun_out/r02_run4_plain.json 2> gpurun_out/r02_run4_plain.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_run4_launches.csv $CMD > gpurun_out/r02_run4_ncu.log 2>&1; echo "ncu rc=$?"
cut -c1-200 gpurun_out/r02_run4_plain.json
