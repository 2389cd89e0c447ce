#!/bin/bash
# quantised 64-byte four-wide records under the compact stack: parity, then C5 on / off (one upload each: the form is chosen at upload)
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q > gpurun_out/r02_run25_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_run25_pytest.log
WRT_QUANT_RECORDS=0 python tools/c5_sweep.py 32 "" "" 2>&1 | tee gpurun_out/r02_run25_q0.txt
WRT_QUANT_RECORDS=1 python tools/c5_sweep.py 32 "" "" "WRT_WF_NODE_SHIFT=1" "WRT_WF_LEAF_BURST=1" 2>&1 | tee gpurun_out/r02_run25_q1.txt
