#!/bin/bash
# launch list (ncu --metrics gpu__time_duration.sum) of the wavefront on C5 at full resolution, one pipeline, 4 spp
set -u
mkdir -p gpurun_out
C5="python bench.py --workload C5 --spp 4 --steps 1 --warmup 0 --no-cpu-baseline --no-all-workloads"
WRT_WF_PIPELINES=1 $C5 > gpurun_out/r02_launches_c5.json 2> gpurun_out/r02_launches_c5.err && WRT_WF_PIPELINES=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches_c5.csv $C5 > gpurun_out/ncu_launches.log 2>&1; echo "ncu rc=$?"
python tools/ncu_launches.py gpurun_out/r02_launches_c5.csv | head -20
