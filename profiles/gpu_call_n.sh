#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time BIC_B200_LIB=$PWD/binary-image-compression_b200/libbic_b200_dbg.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_pipeline.py -m gpu -x -q -k "golomb or encode or pipeline" ) > gpurun_out/n_pytest_dbg.log 2>&1
echo "pytest(debug checks) rc=$? $(tail -4 gpurun_out/n_pytest_dbg.log | head -1)"
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/n_pytest.log 2>&1
echo "pytest rc=$? $(tail -4 gpurun_out/n_pytest.log | head -1)"
timeout 600 python profiles/coder_sweep.py 31 0.001,0.003,0.01,0.03 > gpurun_out/n_coder_sweep.json 2> gpurun_out/n_coder_sweep.err
echo "sweep rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/n_coder_sweep.json'):
    d=json.loads(l); print(d["rho"], round(d["encode_ms"],3), round(d["encode_GBps_in"],1), round(d["decode_ms"],3), round(d["decode_GBps_out"],1), d["roundtrip_ok"], d.get("kernel_ms"))
PY
for gl in 1 2; do
( time timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --gol-list $gl ) > gpurun_out/n_bench_gl$gl.json 2> gpurun_out/n_bench_gl$gl.err
echo "bench gol_list=$gl rc=$?"; tail -c 300 gpurun_out/n_bench_gl$gl.err
python - $gl <<'PY'
import json,sys
try:
    l=[x for x in open(f'gpurun_out/n_bench_gl{sys.argv[1]}.json') if x.startswith('{')][0]
    d=json.loads(l)
    print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3), 'launches', d['gpu_launches'])
    for k,v in d['roofline']['per_kernel'].items():
        if 'gol' in k: print('  ',k,v['ms_per_step'],v['launches_per_step'])
except Exception as ex:
    print('no line', ex)
PY
done
