#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_match.py tests/test_gpu_full_size.py tests/test_host_shim.py -m gpu -x -q ) > gpurun_out/b_pytest.log 2>&1
echo "pytest rc=$? $(tail -4 gpurun_out/b_pytest.log | head -1)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pass -c 6 -o gpurun_out/b_microbench profiles/_bin/microbench > gpurun_out/b_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/b_ncu.log
ls -la gpurun_out/b_microbench.ncu-rep
