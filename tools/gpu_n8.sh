#!/bin/bash
# 8-GPU box: the stated C5 config, the C2 headline, the device-group tests and the CLI on all eight devices
set -u
mkdir -p gpurun_out
nvidia-smi -L | wc -l
bash tools/gpu_c5_campaign.sh 8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r02_scale_c2_n8.json 2> gpurun_out/r02_scale_c2_n8.err; echo "c2 n8 rc=$?"
python - <<'P'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02_scale_c2_n8.json').read().splitlines() if l.startswith('{"metric')][-1])
    print('C2 N=8', round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],1), 'ms', 'e2e', round(d['e2e']['value'],1), 'gather', round(d['gather_ms_per_step'],2), d['kernel_ms_per_rank']['min'], d['kernel_ms_per_rank']['max'], d['mean_radiance'])
except Exception as e: print('ERR', e)
P
python -m pytest tests/test_gpu_configs.py -m gpu -x -q -k "group" > gpurun_out/r02_n8_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_n8_pytest.log
./zig-weekend-raytracer_b200/weekend-raytracer --image_width=1024 --image_height=1024 --samples_per_pixel=1000 --ray_bounce_max_depth=50 --scene=cornell_box --devices=0,1,2,3,4,5,6,7 --image_out_path=gpurun_out/cli_8gpu.ppm --writer=device 2>&1 | tail -5
head -c 20 gpurun_out/cli_8gpu.ppm | head -2; md5sum gpurun_out/cli_8gpu.ppm; rm -f gpurun_out/cli_8gpu.ppm
