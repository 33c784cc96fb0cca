#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( BIC_B200_LIB=$PWD/binary-image-compression_b200/libbic_b200_dbg.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "wide_tiles or mixed" ) > gpurun_out/o_pytest_dbg.log 2>&1
echo "pytest(debug) rc=$? $(tail -1 gpurun_out/o_pytest_dbg.log)"
( timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "wide_tiles or mixed" ) > gpurun_out/o_pytest.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/o_pytest.log)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_g2_ -s 8 -c 4 -o gpurun_out/o_g2_sparse -f python profiles/coder_sweep.py 31 0.001 > gpurun_out/o_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/o_ncu.log
ls -la gpurun_out/o_g2_sparse.ncu-rep
