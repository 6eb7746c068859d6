#!/bin/bash
# longest-first class sort (TSIDB_SCHED_HINT) on / off: tick and e2e of the default workload, the replayed rollout and a
# mid-size batch:  gpurun -- 'bash tools/gpu_hint.sh <tag> [tests|notests]'
TAG=${1:-h}; TESTS=${2:-tests}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build_$TAG.log 2>&1 || { echo "build failed"; tail -20 gpurun_out/build_$TAG.log; exit 1; }
if [ "$TESTS" = "tests" ]; then
  timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_$TAG.log 2>&1
  echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log | cut -c1-300
fi
line() { echo "$1: value $(grep -o '"value": [0-9.]*' $2 | head -1 | grep -o '[0-9.]*$') ms $(grep -o '"ms_per_step": [0-9.]*' $2 | head -1 | grep -o '[0-9.]*$') e2e $(grep -o '"e2e": {"value": [0-9.]*' $2 | grep -o '[0-9.]*$') devrefs $(grep -o '"e2e_device_refs": {"value": [0-9.]*' $2 | grep -o '[0-9.]*$')"; }
for hint in 1 0; do
  TSIDB_SCHED_HINT=$hint python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/hint${hint}_$TAG.log 2>&1; line "hint=$hint walking65536" gpurun_out/hint${hint}_$TAG.log
  TSIDB_SCHED_HINT=$hint python bench.py --steps 10 --warmup 3 --no-cpu-baseline --data replay > gpurun_out/hint${hint}_replay_$TAG.log 2>&1; line "hint=$hint replay" gpurun_out/hint${hint}_replay_$TAG.log
  TSIDB_SCHED_HINT=$hint python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --batch 8192 > gpurun_out/hint${hint}_8192_$TAG.log 2>&1; line "hint=$hint batch 8192" gpurun_out/hint${hint}_8192_$TAG.log
done
exit 0
