#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "wavefront or traversals_agree or boundary or gate1" > gpurun_out/r02_run5_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_run5_pytest.log
for eng in megakernel wavefront; do
  python bench.py --workload C5 --engine $eng --spp 16 --steps 1 --warmup 1 --no-cpu-baseline --no-all-workloads > gpurun_out/r02_run5_c5_$eng.json 2> gpurun_out/r02_run5_c5_$eng.err; echo "c5 $eng rc=$?"
  python bench.py --workload C3 --engine $eng --steps 2 --warmup 1 --no-cpu-baseline --no-all-workloads > gpurun_out/r02_run5_c3_$eng.json 2> gpurun_out/r02_run5_c3_$eng.err; echo "c3 $eng rc=$?"
done
python bench.py --spp 256 --steps 2 --warmup 1 --no-cpu-baseline --no-all-workloads > gpurun_out/r02_run5_c2_256.json 2> gpurun_out/r02_run5_c2.err; echo "c2 rc=$?"
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_run5_c*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],1),'ms', 'e2e', round(d['e2e']['value'],1), d['config'].get('engine'), d.get('mean_radiance'))
    except Exception as e: print(f, 'ERR', e)
P
CMD="python bench.py --workload C5 --res 960x540 --spp 16 --engine wavefront --steps 1 --warmup 0 --no-cpu-baseline --no-all-workloads"
$CMD > gpurun_out/r02_run5_plain.json 2> gpurun_out/r02_run5_plain.err && ncu --set full --clock-control none --import-source on -k regex:wf_extend_ordered -s 3 -c 1 -o gpurun_out/r02_run5_wfext $CMD > gpurun_out/r02_run5_ncu.log 2>&1; echo "ncu rc=$?"
