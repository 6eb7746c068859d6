#!/bin/bash
# multi-GPU bench lines on one box:  gpurun --gpus N -- 'bash tools/gpu_multi.sh <tag> "<N list>" "<config list>"'
TAG=${1:-m}; NS=${2:-8}; CFGS=${3:-walking65536}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build_$TAG.log 2>&1 || { echo "build failed"; exit 1; }
nvidia-smi topo -m > gpurun_out/topo_$TAG.log 2>&1; nproc >> gpurun_out/topo_$TAG.log
for N in $NS; do for C in $CFGS; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py \
      --gpus $N --steps 10 --warmup 3 --config $C --no-cpu-baseline > gpurun_out/bench_${TAG}_${C}_n$N.log 2> gpurun_out/bench_${TAG}_${C}_n$N.err
  echo "N=$N $C rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/bench_${TAG}_${C}_n$N.log | head -2 | tr '\n' ' ')"
done; done
exit 0
