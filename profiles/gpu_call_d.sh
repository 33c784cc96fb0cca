#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golomb or encode" ) > gpurun_out/d_pytest.log 2>&1
echo "pytest rc=$? $(tail -4 gpurun_out/d_pytest.log | head -1)"
( time timeout 600 python bench.py --steps 10 --warmup 3 ) > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err
echo "bench rc=$?"; tail -c 800 gpurun_out/d_bench.err
for f in d_bench; do python - "$f" <<'PY'
import json,sys
f=sys.argv[1]
try:
    l=[x for x in open(f'gpurun_out/{f}.json') if x.startswith('{')][0]
    d=json.loads(l)
    print(f, 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3), 'launches', d['gpu_launches'], d['config'].get('pipeline'))
    for k,v in d['roofline']['per_kernel'].items(): print('  ',k,v['ms_per_step'],v['launches_per_step'])
except Exception as ex:
    print(f, 'no line', ex)
PY
done
timeout 600 python profiles/coder_sweep.py 31 > gpurun_out/d_coder_sweep.json 2> gpurun_out/d_coder_sweep.err
echo "sweep rc=$?"; cat gpurun_out/d_coder_sweep.json | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d["rho"], round(d["encode_ms"],3), round(d["encode_GBps_in"],1), round(d["decode_ms"],3), round(d["decode_GBps_out"],1), d["roundtrip_ok"], d.get("kernel_ms"))
"
