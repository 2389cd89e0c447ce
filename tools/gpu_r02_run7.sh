#!/bin/bash
set -u
mkdir -p gpurun_out
python bench.py --workload C3 --steps 2 --warmup 1 --no-cpu-baseline --no-all-workloads > gpurun_out/r02_run7_c3.json 2> gpurun_out/r02_run7_c3.err; echo "c3 rc=$?"
grep -o '"value": [0-9.]*' gpurun_out/r02_run7_c3.json | head -1
CMD="python bench.py --workload C5 --res 960x540 --spp 16 --engine megakernel --steps 1 --warmup 0 --no-cpu-baseline --no-all-workloads"
$CMD > gpurun_out/r02_run7_plain_mk.json 2> gpurun_out/r02_run7_plain_mk.err && ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 0 -c 1 -o gpurun_out/r02_run7_mk $CMD > gpurun_out/r02_run7_ncu_mk.log 2>&1; echo "ncu mk rc=$?"
CMD="python bench.py --workload C5 --res 960x540 --spp 16 --engine wavefront --steps 1 --warmup 0 --no-cpu-baseline --no-all-workloads"
$CMD > gpurun_out/r02_run7_plain_wf.json 2> gpurun_out/r02_run7_plain_wf.err && ncu --set full --clock-control none --import-source on -k regex:wf_extend_ordered -s 3 -c 1 -o gpurun_out/r02_run7_wf $CMD > gpurun_out/r02_run7_ncu_wf.log 2>&1; echo "ncu wf rc=$?"
grep -o '"value": [0-9.]*' gpurun_out/r02_run7_plain_mk.json gpurun_out/r02_run7_plain_wf.json | head -4
