#!/bin/bash
# first GPU call of round 2: re-baseline + new parity tests + microbenchmarks + smoke under the sanitizer
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/a_gpu.txt 2>&1
nproc >> gpurun_out/a_gpu.txt; free -g >> gpurun_out/a_gpu.txt
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/a_pytest.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/a_pytest.log)"
( time python bench.py --steps 10 --warmup 3 ) > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/a_bench.err
profiles/_bin/microbench > gpurun_out/a_microbench.json 2> gpurun_out/a_microbench.err
echo "microbench rc=$?"; cat gpurun_out/a_microbench.json
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  timeout 240 $CS --tool $tool --error-exitcode 86 --print-limit 20 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/a_sanitizer_smoke_$tool.log 2>&1
  echo "sanitizer $tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/a_sanitizer_smoke_$tool.log | tail -1)"
done
