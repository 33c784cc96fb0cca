#!/bin/bash
# 1 GPU: what the sharded code path costs WITHOUT peers (world = 1), and its launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in "--sharded-slots 16" "--sharded-slots 12" "--sharded-slots 24 --sharded-cluster 16"; do
  tag=$(echo "$cfg" | tr -d ' -')
  timeout 200 python bench.py --sharded --sharded-only --steps 5 --warmup 3 $cfg > gpurun_out/l_$tag.json 2> gpurun_out/l_$tag.err
  echo "$cfg rc=$?"
  python - "$tag" <<'PY'
import json,sys
try:
    l=[x for x in open(f'gpurun_out/l_{sys.argv[1]}.json') if x.startswith('{')][0]
    d=json.loads(l)['row_sharded']
    print('  N=1 sharded value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'launches', d['gpu_launches'], 'parity', d.get('parity_checked'))
except Exception as ex:
    print('  no line', ex)
PY
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/l_launches.csv python bench.py --sharded --sharded-only --steps 1 --warmup 3 > gpurun_out/l_ncu.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/l_launches.csv', errors='ignore')) if len(r)>5]
hdr=None; agg=collections.defaultdict(lambda:[0,0.0])
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if not hdr: continue
    d=dict(zip(hdr,r))
    try: v=float(d['Metric Value'].replace(',',''))
    except: continue
    u=d.get('Metric Unit','')
    v = v/1e3 if u in ('ns','nsecond') else (v*1e3 if u in ('ms','msecond') else v)   # -> us
    k=d['Kernel Name'].split('(')[0]
    agg[k][0]+=1; agg[k][1]+=v
tot=sum(v[1] for v in agg.values())
print('total us', round(tot), 'launches', sum(v[0] for v in agg.values()))
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:40]:
    print(f'{v[1]/tot*100:6.2f}%  {v[0]:6d}  {v[1]/v[0]:9.2f} us  {k[:90]}')
PY
