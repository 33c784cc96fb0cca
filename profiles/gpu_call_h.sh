#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time BIC_B200_LIB=$PWD/binary-image-compression_b200/libbic_b200_dbg.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py -m gpu -x -q -k "golomb or encode or pipeline" ) > gpurun_out/h_pytest_dbg.log 2>&1
echo "pytest(debug checks) rc=$? $(tail -4 gpurun_out/h_pytest_dbg.log | head -1)"
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py tests/test_gpu_full_size.py -m gpu -x -q ) > gpurun_out/h_pytest.log 2>&1
echo "pytest rc=$? $(tail -4 gpurun_out/h_pytest.log | head -1)"
( time timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err
echo "bench rc=$?"; tail -c 400 gpurun_out/h_bench.err
python - <<'PY'
import json
try:
    l=[x for x in open('gpurun_out/h_bench.json') if x.startswith('{')][0]
    d=json.loads(l)
    print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3), 'launches', d['gpu_launches'])
    for k,v in d['roofline']['per_kernel'].items():
        if 'gol' in k: print('  ',k,v['ms_per_step'],v['launches_per_step'])
except Exception as ex:
    print('no line', ex)
PY
timeout 600 python profiles/coder_sweep.py 31 > gpurun_out/h_coder_sweep.json 2> gpurun_out/h_coder_sweep.err
echo "sweep rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/h_coder_sweep.json'):
    d=json.loads(l); print(d["rho"], round(d["encode_ms"],3), round(d["encode_GBps_in"],1), round(d["decode_ms"],3), round(d["decode_GBps_out"],1), d["roundtrip_ok"], d.get("kernel_ms"))
PY
