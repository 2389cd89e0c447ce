#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "wavefront or traversals_agree" > gpurun_out/r02_run8_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_run8_pytest.log
for eng in megakernel wavefront; do
  python bench.py --workload C5 --engine $eng --spp 16 --steps 1 --warmup 1 --no-cpu-baseline --no-all-workloads > gpurun_out/r02_run8_c5_$eng.json 2> gpurun_out/r02_run8_c5_$eng.err; echo "c5 $eng rc=$?"
done
python bench.py --workload C5 --engine wavefront --spp 64 --steps 1 --warmup 1 --warmup-spp 4 --no-cpu-baseline --no-all-workloads > gpurun_out/r02_run8_c5_wf64.json 2> gpurun_out/r02_run8_c5_wf64.err; echo "c5 wf64 rc=$?"
python bench.py --workload C3 --engine wavefront --steps 2 --warmup 1 --no-cpu-baseline --no-all-workloads > gpurun_out/r02_run8_c3_wavefront.json 2> gpurun_out/r02_run8_c3_wf.err; echo "c3 wf rc=$?"
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_run8_c*.json')):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith('{"metric')][-1]); print(f, round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],1),'ms', 'e2e', round(d['e2e']['value'],1), d['config'].get('engine'), d.get('mean_radiance'))
    except Exception as e: print(f, 'ERR', e)
P
