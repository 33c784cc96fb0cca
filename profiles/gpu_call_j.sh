#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_pipeline.py tests/test_multi_gpu.py -m gpu -x -q ) > gpurun_out/j_pytest.log 2>&1
echo "pytest rc=$? $(tail -4 gpurun_out/j_pytest.log | head -1)"; tail -25 gpurun_out/j_pytest.log
