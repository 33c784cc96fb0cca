#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time BIC_B200_LIB=$PWD/binary-image-compression_b200/libbic_b200_dbg.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py tests/test_gpu_edges.py tests/test_gpu_full_size.py -m gpu -x -q ) > gpurun_out/g_pytest_dbg.log 2>&1
echo "pytest(debug checks) rc=$? $(tail -4 gpurun_out/g_pytest_dbg.log | head -1)"
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/g_pytest.log 2>&1
echo "pytest rc=$? $(tail -4 gpurun_out/g_pytest.log | head -1)"
( time timeout 600 python bench.py --steps 10 --warmup 3 ) > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/g_bench.err
python - <<'PY'
import json
try:
    l=[x for x in open('gpurun_out/g_bench.json') if x.startswith('{')][0]
    d=json.loads(l)
    print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3), 'launches', d['gpu_launches'], d['config'].get('pipeline'))
    print('copy floor', d.get('e2e_copy_floor')); print('sweep', d.get('coder_sweep'))
    print('roofline', {k:d['roofline'][k] for k in ('bound','kernel','frac','largest_full_grid_kernel')})
    for k,v in d['roofline']['per_kernel'].items(): print('  ',k,v['ms_per_step'],v['launches_per_step'],v.get('bound'),v.get('frac_of_popc_peak'))
except Exception as ex:
    print('no line', ex)
PY
