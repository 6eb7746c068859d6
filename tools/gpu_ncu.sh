#!/bin/bash
# one full ncu capture of the tick kernels of one bench tick:  gpurun -- 'bash tools/gpu_ncu.sh <tag> [kernel regex]'
TAG=${1:-n}; RX=${2:-"tsidb_(activeset|eliminate|dynamics)_kernel"}; SKIP=${3:-24}; COUNT=${4:-7}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build_$TAG.log 2>&1 || { echo "build failed"; tail -20 gpurun_out/build_$TAG.log; exit 1; }
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RX" --launch-skip $SKIP --launch-count $COUNT -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_$TAG.log | cut -c1-200
