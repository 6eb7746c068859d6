#!/bin/bash
# e2e of the chunked host call against the CTA shape / carve-out of the dynamics kernel (can the next chunk's dynamics
# kernel share SMs with the previous chunk's solver kernels?):  gpurun -- 'bash tools/gpu_e2e_overlap.sh'
mkdir -p gpurun_out
run() { # tag, env...
  tag=$1; shift
  env "$@" python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ovl_$tag.log 2>&1
  echo "$tag: value $(grep -o '"value": [0-9.]*' gpurun_out/ovl_$tag.log | head -1 | grep -o '[0-9.]*$') e2e $(grep -o '"e2e": {"value": [0-9.]*' gpurun_out/ovl_$tag.log | grep -o '[0-9.]*$') devrefs $(grep -o '"e2e_device_refs": {"value": [0-9.]*' gpurun_out/ovl_$tag.log | grep -o '[0-9.]*$') $(grep -o '"dynamics": [0-9.]*' gpurun_out/ovl_$tag.log | head -1)"
}
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
run base A=1
run base_carve100 TSIDB_D_CARVEOUT=100
cp tsid_control_b200/csrc/tsidb_const.h /tmp/const.bak
build() { (cd tsid_control_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -o libtsidb.so tsidb.cu) || echo "build failed"; }
sed -i -E 's/^#define TSIDB_WARPS_PER_BLOCK [0-9]+/#define TSIDB_WARPS_PER_BLOCK 8/; s/^#define TSIDB_D_CTAS_PER_SM [0-9]+/#define TSIDB_D_CTAS_PER_SM 2/' tsid_control_b200/csrc/tsidb_const.h
build
run wpb8x2 A=1
run wpb8x2_carve100 TSIDB_D_CARVEOUT=100
sed -i -E 's/^#define TSIDB_A_CTA_WARPS_DS [0-9]+/#define TSIDB_A_CTA_WARPS_DS 2/; s/^#define TSIDB_A_CTA_WARPS_SS [0-9]+/#define TSIDB_A_CTA_WARPS_SS 2/' tsid_control_b200/csrc/tsidb_const.h
build
run wpb8x2_a2_carve100 TSIDB_D_CARVEOUT=100
cp /tmp/const.bak tsid_control_b200/csrc/tsidb_const.h
