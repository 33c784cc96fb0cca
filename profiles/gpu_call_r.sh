#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 30 python profiles/coder_sweep.py 31 0.001,0.01,0.2 > gpurun_out/r_sweep.json 2> gpurun_out/r_sweep.err
BIC_B200_LIB=$PWD/binary-image-compression_b200/libbic_b200_alt.so timeout 20 python profiles/coder_sweep.py 31 0.2 > gpurun_out/r_sweep_alt.json 2> gpurun_out/r_sweep_alt.err
python - <<'PY'
import json
for f in ('r_sweep','r_sweep_alt'):
    for l in open(f'gpurun_out/{f}.json'):
        d=json.loads(l); print(f, d["rho"], round(d["encode_ms"],3), d["roundtrip_ok"], d["kernel_ms"])
PY
( timeout 70 python -m pytest tests -m gpu -x -q ) > gpurun_out/r_pytest.log 2>&1
echo "pytest rc=$? $(tail -3 gpurun_out/r_pytest.log | grep -E 'passed|failed')"
