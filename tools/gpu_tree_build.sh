#!/bin/bash
# device tree build against the host build (bytes, frames, timing)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_build.py -m gpu -x -q -s > gpurun_out/r02_build_pytest.log 2>&1; echo "pytest rc=$?"
grep -h "records, host\|2^20 primitives\|passed\|failed\|Error\|error" gpurun_out/r02_build_pytest.log | head -40
