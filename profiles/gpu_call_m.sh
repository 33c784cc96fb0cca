#!/bin/bash
# 8 GPUs: sharded pipeline, slots / cluster sweep (sharded arm only)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in "--sharded-slots 12" "--sharded-slots 8" "--sharded-slots 12 --sharded-cluster 4"; do
  tag=$(echo "$cfg" | tr -d ' -')
  ( time timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 5 --warmup 3 --sharded-only $cfg ) > gpurun_out/m_$tag.json 2> gpurun_out/m_$tag.err
  echo "$cfg rc=$?"
  python - "$tag" <<'PY'
import json,sys
try:
    l=[x for x in open(f'gpurun_out/m_{sys.argv[1]}.json') if x.startswith('{')][0]
    d=json.loads(l)['row_sharded']
    print('  N=8 sharded value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'launches', d['gpu_launches'], 'parity', d.get('parity_checked'), 'numa', d.get('numa'))
except Exception as ex:
    print('  no line', ex)
PY
  grep real gpurun_out/m_$tag.err
done
