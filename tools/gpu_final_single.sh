#!/bin/bash
# final single-GPU evidence of round 2 (after the device tree build, 256-bit loads, reordering, 8 blocks/SM): all GPU tests,
# smoke, the bench exactly as the driver runs it, the reference arm, the stated C5 on one GPU, a launch list of the default bench
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/r02_final2_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_final2_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_final2_smoke.log
python bench.py > gpurun_out/r02_bench_n1_default.json 2> gpurun_out/r02_bench_n1_default.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; echo "ref rc=$?"
python bench.py --workload C5 --fused-e2e --warmup-spp 4 --steps 1 --warmup 3 --no-all-workloads > gpurun_out/r02_c5_full_n1.json 2> gpurun_out/r02_c5_full_n1.err; echo "c5 rc=$?"
python - <<'P'
import json
def line(p, key='{"metric'):
    return json.loads([l for l in open(p).read().splitlines() if l.startswith(key)][-1])
d=line('gpurun_out/r02_bench_n1_default.json')
print('C2', round(d['value'],1), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],1), 'issue frac', round(d['roofline_issue']['frac'],4), 'cpu', d['cpu_baseline']['value'])
for k,v in d.get('workloads',{}).items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('upload_ms'), v.get('error'))
r=line('gpurun_out/r02_bench_reference_arm.json','{"impl'); print('reference arm', r['value'])
c=line('gpurun_out/r02_c5_full_n1.json'); print('C5 full', round(c['value'],1), round(c['ms_per_step']/1e3,2), 's e2e', round(c['e2e']['value'],1), c['mean_radiance'], c.get('cpu_baseline',{}).get('value'))
P
