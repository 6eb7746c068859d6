#!/bin/bash
# quick GPU session: GPU parity tests (optional), the bench line without the CPU legs, then build variants
#   gpurun --timeout 1500 -- 'bash tools/gpu_quick.sh <tag> [tests|notests] "VARIANT VARIANT ..."'
TAG=${1:-q}; TESTS=${2:-tests}; VARS=$3
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build_$TAG.log 2>&1 || { echo "build failed"; tail -20 gpurun_out/build_$TAG.log; exit 1; }
if [ "$TESTS" = "tests" ]; then
  timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_$TAG.log 2>&1
  echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log | cut -c1-300
fi
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err
echo "bench rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/bench_$TAG.log | head -1) $(grep -o '"kernel_ms_all": {[^}]*}' gpurun_out/bench_$TAG.log)"
if [ -n "$VARS" ]; then bash tools/gpu_variants.sh "$VARS"; fi
exit 0
