#!/bin/bash
# compute-sanitizer passes over the hot path (race detection / sanitizers). Run on a GPU box:
#   bash profiles/sanitize.sh            -> gpurun_out/sanitizer_*.log (+ a summary on stdout)
# memcheck + racecheck + synccheck + initcheck over smoke() and over the tests that drive the kernels with shared-memory
# atomics, DSMEM reductions and rotating buffers (the cluster chain, the bucket fill, the Golomb walks).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
SMOKE='import __graft_entry__ as g; g.smoke()'
CHAIN='tests/test_gpu_parity.py::test_cluster_chain_any_cluster_size tests/test_gpu_parity.py::test_dense_coefficients_many_shared_rows'
CODER='tests/test_gpu_parity.py::test_golomb_stream_is_byte_identical_and_decodes'
run() {  # name, tool, timeout, command...
  local name=$1 tool=$2 to=$3; shift 3
  timeout "$to" $CS --tool "$tool" --error-exitcode 86 --print-limit 20 "$@" > "gpurun_out/sanitizer_${name}.log" 2>&1
  local rc=$?
  echo "== $name ($tool) rc=$rc: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_${name}.log | tail -1)"
}
run smoke_memcheck memcheck 300 python -c "$SMOKE"
run smoke_racecheck racecheck 300 python -c "$SMOKE"
run smoke_synccheck synccheck 300 python -c "$SMOKE"
run smoke_initcheck initcheck 300 python -c "$SMOKE"
run chain_memcheck memcheck 600 python -m pytest -x -q $CHAIN -k "8 or shared"
run chain_racecheck racecheck 900 python -m pytest -x -q $CHAIN -k "8--1 or shared"
run chain_synccheck synccheck 600 python -m pytest -x -q $CHAIN -k "8--1 or shared"
run coder_memcheck memcheck 600 python -m pytest -x -q $CODER -k "0-300 or 0-2000 or 1-2000 or 0-1-200000"
run coder_racecheck racecheck 600 python -m pytest -x -q $CODER -k "0-300 or 0-2000"
