#!/bin/bash
set -u
mkdir -p gpurun_out
WRT_TRACE_BUILD=1 python bench.py --steps 1 --warmup 1 > gpurun_out/r02_trace2.json 2> gpurun_out/r02_trace2.err; echo "rc=$?"
grep "wrt trace" gpurun_out/r02_trace2.err | tail -14
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r02_trace2.json').read().splitlines() if l.startswith('{"metric')][-1])
print({k:(round(v['upload_ms'],1), round(v['value'],1)) for k,v in d['workloads'].items()})
P
