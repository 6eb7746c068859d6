#!/bin/bash
# One GPU-box session: GPU tests, the bench line, the launch list and a full ncu capture of one tick.
#   gpurun --timeout 1800 -- 'bash tools/gpu_round.sh r5a [tests|notests] [ncu|noncu] [configs|noconfigs]'
# Everything lands in gpurun_out/ (the only directory that travels back).
TAG=${1:-rX}; TESTS=${2:-tests}; NCU=${3:-ncu}; CONFIGS=${4:-noconfigs}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build_$TAG.log 2>&1 || { echo "build failed"; tail -20 gpurun_out/build_$TAG.log; exit 1; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.log 2>&1
nproc >> gpurun_out/smi_$TAG.log
if [ "$TESTS" = "tests" ]; then
  timeout 1200 python -m pytest tests -q -m gpu -s > gpurun_out/pytest_$TAG.log 2>&1
  echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log | cut -c1-300
fi
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"; tail -c 600 gpurun_out/bench_$TAG.log
if [ "$CONFIGS" = "configs" ]; then
  for c in standing4096 legacy16384 mixed1M; do
    timeout 900 python bench.py --steps 10 --warmup 3 --config $c > gpurun_out/bench_${TAG}_$c.log 2> gpurun_out/bench_${TAG}_$c.err
    echo "bench $c rc=$?"
  done
  timeout 600 python bench.py --steps 20 --warmup 3 --data replay > gpurun_out/bench_${TAG}_replay.log 2> gpurun_out/bench_${TAG}_replay.err
  echo "bench replay rc=$?"
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference.log 2> gpurun_out/bench_${TAG}_reference.err
  echo "bench reference rc=$?"
fi
if [ "$NCU" = "ncu" ]; then
  timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/plain_$TAG.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tsidb_(activeset|eliminate|dynamics)_kernel' \
      --launch-skip 24 --launch-count 7 -f -o gpurun_out/prof_$TAG python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$TAG.log 2>&1
  echo "ncu rc=$?"; tail -2 gpurun_out/ncu_$TAG.log | cut -c1-200
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 20 --launch-count 60 --csv \
      --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu2_$TAG.log 2>&1
fi
exit 0
