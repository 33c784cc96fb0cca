#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_parity.py -m gpu -x -q ) > gpurun_out/c_pytest.log 2>&1
echo "pytest rc=$? $(tail -4 gpurun_out/c_pytest.log | head -1)"
( time timeout 600 python bench.py --steps 10 --warmup 3 ) > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/c_bench.err
( time timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --first-batch 3 --next-batch 2 ) > gpurun_out/c_bench_b3.json 2> gpurun_out/c_bench_b3.err
echo "bench b3 rc=$?"
( time timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --streams 32 ) > gpurun_out/c_bench_s32.json 2> gpurun_out/c_bench_s32.err
echo "bench s32 rc=$?"
( time timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --pool ) > gpurun_out/c_bench_pool.json 2> gpurun_out/c_bench_pool.err
echo "bench pool rc=$?"
for f in c_bench c_bench_b3 c_bench_s32 c_bench_pool; do python - "$f" <<'PY'
import json,sys
f=sys.argv[1]
try:
    l=[x for x in open(f'gpurun_out/{f}.json') if x.startswith('{')][0]
    d=json.loads(l)
    print(f, 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3), 'launches', d['gpu_launches'], d['config'].get('pipeline'))
except Exception as ex:
    print(f, 'no line', ex)
PY
done
