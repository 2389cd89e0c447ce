#!/bin/bash
# C5 on 8 GPUs with 32 sample chunks per pixel instead of the library's 17 (finer jobs keep each device's pool full)
set -u
mkdir -p gpurun_out
ARGS="--workload C5 --gpus 8 --steps 1 --warmup 3 --warmup-spp 4 --fused-e2e --no-all-workloads --chunks 32"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py $ARGS > gpurun_out/r02_c5_full_n8_chunks32.json 2> gpurun_out/r02_c5_full_n8_chunks32.err
echo "rc=$?"
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r02_c5_full_n8_chunks32.json').read().splitlines() if l.startswith('{"metric')][-1])
print('C5 full N=8 chunks 32', round(d['value'],1), 'Mrays/s', round(d['ms_per_step']/1e3,2), 's/frame', 'e2e', round(d['e2e']['value'],1), 'gather ms', round(d['gather_ms_per_step'],2), d['kernel_ms_per_rank']['min'], d['kernel_ms_per_rank']['max'], 'mean', d['mean_radiance'])
P
