#!/bin/bash
# usage: tools/gpu_variants.sh "NAME=VAL[,NAME=VAL...] ..."  — per variant: patch the #defines in tsidb_const.h / tsidb_kernels.cuh,
# rebuild libtsidb.so on the GPU box, print the bench value and per-kernel times
mkdir -p gpurun_out
cp tsid_control_b200/csrc/tsidb_const.h /tmp/const.bak; cp tsid_control_b200/csrc/tsidb_kernels.cuh /tmp/kern.bak
for v in $1; do
  cp /tmp/const.bak tsid_control_b200/csrc/tsidb_const.h; cp /tmp/kern.bak tsid_control_b200/csrc/tsidb_kernels.cuh
  for kv in ${v//,/ }; do
    k=${kv%%=*}; val=${kv##*=}
    sed -i -E "s/^#define $k [0-9]+/#define $k $val/" tsid_control_b200/csrc/tsidb_const.h tsid_control_b200/csrc/tsidb_kernels.cuh
  done
  (cd tsid_control_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -o libtsidb.so tsidb.cu) || { echo "build failed $v"; continue; }
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/var_$v.log 2>&1
  echo "variant $v: $(grep -o '"value": [0-9.]*' gpurun_out/var_$v.log | head -1) e2e $(grep -o '"e2e": {"value": [0-9.]*' gpurun_out/var_$v.log | grep -o '[0-9.]*$') $(grep -o '"kernel_ms_all": {[^}]*}' gpurun_out/var_$v.log)"
  for b in $VAR_BATCHES; do   # optional: the tick alone at other batch sizes
    python bench.py --batch $b --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/var_${v}_$b.log 2>&1
    echo "  batch $b: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/var_${v}_$b.log | head -1) $(grep -o '"kernel_ms_all": {[^}]*}' gpurun_out/var_${v}_$b.log)"
  done
done
cp /tmp/const.bak tsid_control_b200/csrc/tsidb_const.h; cp /tmp/kern.bak tsid_control_b200/csrc/tsidb_kernels.cuh
