#!/bin/bash
# representative profiles of the per-lane kernel: long sample chunks (256 spp) on a quarter-size frame
set -u
mkdir -p gpurun_out
C5="python bench.py --workload C5 --res 960x540 --spp 256 --steps 1 --warmup 1 --no-cpu-baseline"
$C5 > gpurun_out/r02_c5_q256.json 2> gpurun_out/r02_c5_q256.err && ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -o gpurun_out/r02_c5_q256 $C5 > gpurun_out/ncu_c5b.log 2>&1; echo "ncu c5 rc=$?"
C3="python bench.py --workload C3 --res 960x540 --spp 512 --steps 1 --warmup 1 --no-cpu-baseline"
$C3 > gpurun_out/r02_c3_q512.json 2> gpurun_out/r02_c3_q512.err && ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -o gpurun_out/r02_c3_q512 $C3 > gpurun_out/ncu_c3b.log 2>&1; echo "ncu c3 rc=$?"
python bench.py --workload C5 --spp 64 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r02_base_c5_spp64.json 2>> gpurun_out/r02_c5_q256.err; echo "c5 spp64 rc=$?"
cut -c1-300 gpurun_out/r02_c5_q256.json gpurun_out/r02_c3_q512.json gpurun_out/r02_base_c5_spp64.json
