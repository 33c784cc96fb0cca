#!/bin/bash
# 2 GPUs: sharded parity tests + the N=2 bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/e_gpus.txt; nproc >> gpurun_out/e_gpus.txt
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 4 --warmup 3 --no-cpu-baseline ) > gpurun_out/e_bench8.json 2> gpurun_out/e_bench8.err
echo "bench2 rc=$?"; tail -c 1500 gpurun_out/e_bench8.err
python - <<'PY'
import json
try:
    l=[x for x in open('gpurun_out/e_bench8.json') if x.startswith('{')][0]
    d=json.loads(l)
    print('N=8 value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'parity', d.get('parity_checked'), d.get('parity_planes'))
    print('replicas', d.get('replicas',{}).get('value'), d.get('replicas',{}).get('ms_per_step'))
    rs=d.get('row_sharded',{}); print({k:rs.get(k) for k in ('value','ms_per_step','collectives_per_step','gpu_launches','planes_in_flight_per_rank','iterations_per_plane')})
except Exception as ex:
    print('no line', ex)
PY
