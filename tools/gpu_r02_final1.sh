#!/bin/bash
# final single-GPU evidence: all GPU tests (-s: the tests print their measurements), the bench exactly as the driver runs it,
# the reference arm
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/r02_final_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_final_pytest.log
grep -h "rtw_final 64x64\|RMSE vs 2048" gpurun_out/r02_final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02_final_smoke.log
python bench.py > gpurun_out/r02_bench_n1_default.json 2> gpurun_out/r02_bench_n1_default.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; echo "ref rc=$?"
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_n1_default.json').read().splitlines() if l.startswith('{"metric')][-1])
print('C2', round(d['value'],1), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],1), 'issue frac', round(d['roofline_issue']['frac'],4), 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['build'])
for k,v in d.get('workloads',{}).items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('error'))
r=json.loads([l for l in open('gpurun_out/r02_bench_reference_arm.json').read().splitlines() if l.startswith('{"impl')][-1]); print('reference arm', r['value'], r['cpu_baseline'])
P
