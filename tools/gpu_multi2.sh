#!/bin/bash
# 2-GPU box: device groups and the sharded render through the C ABI (tests, bench under torchrun, CLI --devices)
set -u
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_gpu_configs.py -m gpu -x -q -s -k "group or contexts" > gpurun_out/r02_multi2_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_multi2_pytest.log
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --spp 512 --steps 3 --warmup 2 > gpurun_out/r02_multi2_bench.json 2> gpurun_out/r02_multi2_bench.err; echo "bench2 rc=$?"
grep -c "nranks 2" gpurun_out/r02_multi2_bench.err
tail -1 gpurun_out/r02_multi2_bench.json | cut -c1-1500
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --spp 512 --steps 2 --warmup 1 --shard samples > gpurun_out/r02_multi2_bench_samples.json 2> gpurun_out/r02_multi2_bench_samples.err; echo "bench2 samples rc=$?"
tail -1 gpurun_out/r02_multi2_bench_samples.json | cut -c1-400
python bench.py --gpus 1 --spp 512 --steps 3 --warmup 2 --no-cpu-baseline --no-all-workloads > gpurun_out/r02_multi2_bench1.json 2>/dev/null; tail -1 gpurun_out/r02_multi2_bench1.json | cut -c1-300
./zig-weekend-raytracer_b200/weekend-raytracer --image_width=512 --image_height=512 --samples_per_pixel=64 --ray_bounce_max_depth=20 --scene=cornell_box --devices=0,1 --image_out_path=gpurun_out/cli_2gpu.ppm --writer=device 2>&1 | tail -6
./zig-weekend-raytracer_b200/weekend-raytracer --image_width=512 --image_height=512 --samples_per_pixel=64 --ray_bounce_max_depth=20 --scene=cornell_box --device=0 --image_out_path=gpurun_out/cli_1gpu.ppm 2>&1 | tail -4
cmp gpurun_out/cli_1gpu.ppm gpurun_out/cli_2gpu.ppm && echo "CLI frames identical on 1 and 2 GPUs"; rm -f gpurun_out/cli_*.ppm
