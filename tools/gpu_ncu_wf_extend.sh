#!/bin/bash
# ncu --set full of the persistent extend kernel (current build) on C5, one mid-frame launch
set -u
mkdir -p gpurun_out
C5="python bench.py --workload C5 --res 1920x1080 --spp 32 --steps 1 --warmup 0 --no-cpu-baseline --no-all-workloads"
WRT_WF_PIPELINES=1 $C5 > gpurun_out/r02_prof3_c5.json 2> gpurun_out/r02_prof3_c5.err && WRT_WF_PIPELINES=1 ncu --set full --clock-control none --import-source on -k regex:wf_extend_ordered -s 12 -c 1 -o gpurun_out/r02_prof3_wf $C5 > gpurun_out/ncu_prof3.log 2>&1; echo "ncu rc=$?"
cut -c1-200 gpurun_out/r02_prof3_c5.json; tail -3 gpurun_out/ncu_prof3.log
