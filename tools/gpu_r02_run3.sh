#!/bin/bash
# ABI v2 on the GPU: all GPU tests, then the persistent wavefront extend against the megakernel on the per-lane workloads
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/r02_run3_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_run3_pytest.log
for eng in megakernel wavefront; do
  python bench.py --workload C3 --engine $eng --steps 2 --warmup 1 --no-cpu-baseline --no-all-workloads > gpurun_out/r02_run3_c3_$eng.json 2> gpurun_out/r02_run3_c3_$eng.err; echo "c3 $eng rc=$?"
  python bench.py --workload C5 --engine $eng --spp 16 --steps 1 --warmup 1 --no-cpu-baseline --no-all-workloads > gpurun_out/r02_run3_c5_$eng.json 2> gpurun_out/r02_run3_c5_$eng.err; echo "c5 $eng rc=$?"
done
python bench.py --spp 256 --steps 2 --warmup 1 --no-cpu-baseline --no-all-workloads > gpurun_out/r02_run3_c2_256.json 2> gpurun_out/r02_run3_c2.err; echo "c2 rc=$?"
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_run3_c*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],1),'ms', 'e2e', round(d['e2e']['value'],1), d['config'].get('engine'), d.get('mean_radiance'))
    except Exception as e: print(f, 'ERR', e)
P
tail -3 gpurun_out/r02_run3_c*.err
