#!/bin/bash
# round-2 baseline: GPU tests, bench lines of the per-lane workloads, ncu --set full of render_kernel<0,0> on C5 and C3
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_base_pytest.log 2>&1; echo "pytest rc=$?"
python bench.py --workload C3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_base_c3.json 2> gpurun_out/r02_base_c3.err; echo "c3 rc=$?"
python bench.py --workload C5 --spp 16 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_base_c5_spp16.json 2> gpurun_out/r02_base_c5.err; echo "c5 rc=$?"
C5="python bench.py --workload C5 --spp 2 --steps 1 --warmup 1 --no-cpu-baseline"
$C5 > gpurun_out/plain_c5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -o gpurun_out/r02_base_c5 $C5 > gpurun_out/ncu_c5.log 2>&1; echo "ncu c5 rc=$?"
C3="python bench.py --workload C3 --spp 16 --steps 1 --warmup 1 --no-cpu-baseline"
$C3 > gpurun_out/plain_c3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -o gpurun_out/r02_base_c3 $C3 > gpurun_out/ncu_c3.log 2>&1; echo "ncu c3 rc=$?"
tail -3 gpurun_out/r02_base_pytest.log
cat gpurun_out/r02_base_c3.json gpurun_out/r02_base_c5_spp16.json | cut -c1-400
