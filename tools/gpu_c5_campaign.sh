#!/bin/bash
# BASELINE.json configs[4] as stated: 2^20 primitives, 3840x2160, 1 024 spp, depth 20, on N GPUs of this box.
# usage: tools/gpu_c5_campaign.sh N   (N = 1 launches plain python, N > 1 torchrun; one timed frame, fused e2e timing)
set -u
N=${1:-1}
mkdir -p gpurun_out
ARGS="--workload C5 --gpus $N --steps 1 --warmup 3 --warmup-spp 4 --fused-e2e --no-all-workloads"
if [ "$N" = 1 ]; then
  python bench.py $ARGS > gpurun_out/r02_c5_full_n$N.json 2> gpurun_out/r02_c5_full_n$N.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py $ARGS > gpurun_out/r02_c5_full_n$N.json 2> gpurun_out/r02_c5_full_n$N.err
fi
echo "c5 n=$N rc=$?"
python - <<P
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02_c5_full_n$N.json').read().splitlines() if l.startswith('{"metric')][-1])
    print('C5 full N=$N', round(d['value'],1), 'Mrays/s', round(d['ms_per_step']/1e3,2), 's/frame', 'e2e', round(d['e2e']['value'],1), 'gather ms', round(d['gather_ms_per_step'],2), 'kernel ms/rank', d['kernel_ms_per_rank']['min'], d['kernel_ms_per_rank']['max'], 'hbm frac', round(d['roofline']['frac'],4), 'mean', d['mean_radiance'], d.get('cpu_baseline',{}).get('value'))
except Exception as e: print('ERR', e)
P
tail -2 gpurun_out/r02_c5_full_n$N.err | cut -c1-300
