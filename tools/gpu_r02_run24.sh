#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "wavefront or engine or synthetic" > gpurun_out/r02_run24_pytest.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/r02_run24_pytest.log
python tools/c5_sweep.py 32 "" "WRT_WF_NODE_SHIFT=1" "WRT_WF_NODE_SHIFT=3" "WRT_WF_LEAF_BURST=2" "WRT_WF_LEAF_BURST=8" "WRT_WF_LEAF_BURST=1" "WRT_WF_NODE_BURST=8" "WRT_WF_NODE_BURST=64" "WRT_WF_NODE_SHIFT=3 WRT_WF_LEAF_BURST=2" "WRT_WF_PIPELINES=2" "WRT_WF_PIPELINES=6" "" 2>&1 | tee gpurun_out/r02_run24_sweep.txt
