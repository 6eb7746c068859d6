#!/bin/bash
# GPU session for the small-batch path: GPU tests (optional), then the latency probe with the hand-off images in shared
# memory (default) and in global memory (TSIDB_SMALL_LOCAL_N=0):  gpurun -- 'bash tools/gpu_latency.sh <tag> [tests|notests]'
TAG=${1:-lat}; TESTS=${2:-tests}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build_$TAG.log 2>&1 || { echo "build failed"; tail -20 gpurun_out/build_$TAG.log; exit 1; }
if [ "$TESTS" = "tests" ]; then
  timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_$TAG.log 2>&1
  echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log | cut -c1-300
fi
for loc in "" 0; do
  echo "TSIDB_SMALL_LOCAL_N=${loc:-default}" | tee -a gpurun_out/latency_$TAG.log
  PROBE_N=${PROBE_N:-1,32,256} PROBE_SMALL=1024 TSIDB_SMALL_LOCAL_N=$loc timeout 300 python tools/latency_probe.py 2>&1 | tee -a gpurun_out/latency_$TAG.log
done
exit 0
