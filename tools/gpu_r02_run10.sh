#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_run10_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_run10_pytest.log
for eng in auto megakernel; do
  python bench.py --workload C5 --engine $eng --spp 64 --steps 1 --warmup 1 --warmup-spp 4 --fused-e2e --no-cpu-baseline --no-all-workloads > gpurun_out/r02_run10_c5_$eng.json 2> gpurun_out/r02_run10_c5_$eng.err; echo "c5 $eng rc=$?"
done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_run10_c*.json')):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith('{"metric')][-1]); print(f, round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],1),'ms', 'e2e', round(d['e2e']['value'],1), d['config'].get('engine'), d.get('mean_radiance'))
    except Exception as e: print(f, 'ERR', e)
P
