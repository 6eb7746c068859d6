#!/bin/bash
# usage: tools/gpu_variants.sh "<D_WARPS>:<E_WARPS> ..."   — rebuild libtsidb.so on the GPU box per variant and print the bench value
mkdir -p gpurun_out
cp tsid_control_b200/csrc/tsidb_const.h /tmp/const.bak
for v in $1; do
  d=${v%%:*}; e=${v##*:}
  sed -e "s/#define TSIDB_WARPS_PER_BLOCK [0-9]*/#define TSIDB_WARPS_PER_BLOCK $d/" -e "s/#define TSIDB_E_WARPS [0-9]*/#define TSIDB_E_WARPS $e/" /tmp/const.bak > tsid_control_b200/csrc/tsidb_const.h
  (cd tsid_control_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -o libtsidb.so tsidb.cu) || { echo "build failed $v"; continue; }
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/var_${d}_${e}.log 2>&1
  echo "variant D=$d E=$e: $(grep -o '"value": [0-9.]*' gpurun_out/var_${d}_${e}.log | head -1)  $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/var_${d}_${e}.log | head -1)"
done
cp /tmp/const.bak tsid_control_b200/csrc/tsidb_const.h
