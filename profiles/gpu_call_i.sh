#!/bin/bash
# 2 GPUs: sharded tests (new sharded coder) + single-GPU shard tests + N=2 bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_multi_gpu.py tests/test_gpu_full_size.py -m gpu -x -q -k "sharded or shard" ) > gpurun_out/i_pytest.log 2>&1
echo "pytest rc=$? $(tail -4 gpurun_out/i_pytest.log | head -1)"
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline ) > gpurun_out/i_bench2.json 2> gpurun_out/i_bench2.err
echo "bench2 rc=$?"; tail -c 600 gpurun_out/i_bench2.err
python - <<'PY'
import json
try:
    l=[x for x in open('gpurun_out/i_bench2.json') if x.startswith('{')][0]
    d=json.loads(l)
    print('N=2 value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'parity', d.get('parity_checked'), d.get('parity_planes'))
    print('replicas', d.get('replicas',{}).get('value'), d.get('replicas',{}).get('ms_per_step'))
except Exception as ex:
    print('no line', ex)
PY
