#!/bin/bash
# usage (on the GPU box, via gpurun): tools/gpu_check.sh <tag> [full]
# GPU parity tests, bench line, ncu launch list; with "full" also one ncu --set full capture of the tick kernels.
tag=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${tag}.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_${tag}.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${tag}.log 2> gpurun_out/bench_${tag}.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_${tag}.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_${tag}.log 2>&1; echo "ncu launches rc=$?"
if [ "$2" = "full" ]; then
  # one launch of each tick kernel (dynamics + 3 classes x elimination, basis, active set) of the second warm-up tick
  ncu --set full --clock-control none --import-source on -k regex:"tsidb_(dynamics|eliminate|j2|activeset)" -s 10 -c 10 -o gpurun_out/prof_${tag} -f \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${tag}.log 2>&1; echo "ncu full rc=$?"
fi
