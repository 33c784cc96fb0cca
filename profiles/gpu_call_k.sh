#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in "--sharded-slots 24" "--sharded-slots 24 --sharded-cluster 4" "--sharded-slots 12"; do
  tag=$(echo "$cfg" | tr -d ' -')
  ( time timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline $cfg ) > gpurun_out/k_$tag.json 2> gpurun_out/k_$tag.err
  echo "$cfg rc=$?"
  python - "$tag" <<'PY'
import json,sys
try:
    l=[x for x in open(f'gpurun_out/k_{sys.argv[1]}.json') if x.startswith('{')][0]
    d=json.loads(l)
    print('  N=2 value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'parity', d.get('parity_checked'), 'replicas ms', round(d['replicas']['ms_per_step'],3))
except Exception as ex:
    print('  no line', ex)
PY
done
