#!/bin/bash
# final single-GPU verification of the round: full GPU suite (normal + debug-check build), coder sweep A/B, the bench line, launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/p_pytest.log 2>&1
echo "pytest rc=$? $(tail -5 gpurun_out/p_pytest.log | grep -E 'passed|failed')"
( time BIC_B200_LIB=$PWD/binary-image-compression_b200/libbic_b200_dbg.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_pipeline.py tests/test_gpu_match.py -m gpu -x -q -k "golomb or encode or pipeline or match or chain" ) > gpurun_out/p_pytest_dbg.log 2>&1
echo "pytest(debug checks) rc=$? $(tail -5 gpurun_out/p_pytest_dbg.log | grep -E 'passed|failed')"
timeout 300 python profiles/coder_sweep.py 31 > gpurun_out/p_coder_sweep.json 2> gpurun_out/p_coder_sweep.err
echo "sweep rc=$?"
BIC_SWEEP_OPTS=gol_scan=0 timeout 200 python profiles/coder_sweep.py 31 0.001,0.01,0.1 > gpurun_out/p_coder_sweep_scan0.json 2> gpurun_out/p_coder_sweep_scan0.err
BIC_SWEEP_OPTS=gol_scan=0,gol_list=0 timeout 200 python profiles/coder_sweep.py 31 0.001,0.01 > gpurun_out/p_coder_sweep_scan0_list0.json 2> gpurun_out/p_coder_sweep_scan0_list0.err
python - <<'PY'
import json
for f in ('p_coder_sweep','p_coder_sweep_scan0','p_coder_sweep_scan0_list0'):
    print(f)
    for l in open(f'gpurun_out/{f}.json'):
        d=json.loads(l); print(' ', d["rho"], round(d["encode_ms"],3), round(d["encode_GBps_in"],1), round(d["decode_ms"],3), round(d["decode_GBps_out"],1), d["roundtrip_ok"], d.get("kernel_ms"))
PY
( time timeout 600 python bench.py ) > gpurun_out/p_bench.json 2> gpurun_out/p_bench.err
echo "bench rc=$?"; tail -c 300 gpurun_out/p_bench.err
( timeout 300 python bench.py --no-cpu-baseline --gol-scan 2 ) > gpurun_out/p_bench_scan2.json 2> gpurun_out/p_bench_scan2.err
echo "bench(gol_scan 2) rc=$?"
python - <<'PY'
import json
for f in ('p_bench','p_bench_scan2'):
    try:
        l=[x for x in open(f'gpurun_out/{f}.json') if x.startswith('{')][0]
        d=json.loads(l)
        print(f,'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3), 'launches', d['gpu_launches'], 'parity', d.get('parity_checked'), d.get('parity_planes'), 'roofline', d['roofline'].get('kernel'), d['roofline'].get('frac'))
        for k,v in d['roofline']['per_kernel'].items():
            print('  ',k,v.get('ms_per_step'),v.get('launches_per_step'),v.get('bound'),v.get('frac_of_popc_peak'),v.get('frac_of_hbm_peak'))
        print('  cpu', d.get('cpu_baseline'))
    except Exception as ex:
        print(f,'no line', ex)
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/p_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/p_ncu.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/p_launches.csv', errors='ignore')) if len(r)>5]
hdr=None; agg=collections.defaultdict(lambda:[0,0.0])
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if not hdr: continue
    d=dict(zip(hdr,r))
    try: v=float(d['Metric Value'].replace(',',''))
    except: continue
    u=d.get('Metric Unit','')
    v = v/1e3 if u in ('ns','nsecond') else (v*1e3 if u in ('ms','msecond') else v)
    k=d['Kernel Name'].split('(')[0]
    if k.startswith('void at::') or k.startswith('at::'): continue
    agg[k][0]+=1; agg[k][1]+=v
tot=sum(v[1] for v in agg.values())
print('total us (own kernels)', round(tot), 'launches', sum(v[0] for v in agg.values()))
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:30]:
    print(f'{v[1]/tot*100:6.2f}%  {v[0]:6d}  {v[1]/v[0]:9.2f} us  {k[:90]}')
PY
