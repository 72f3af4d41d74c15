#!/bin/bash
# Build an experimental variant of libbpt.so with extra nvcc flags (A/B runs select it with BPT_LIBRARY=...).
#   tools/build_variant.sh NAME "-DBPT_TRACE_MIN_CTAS=8 ..."   ->  buas_pathtracer_b200/csrc/build/variants/libbpt_NAME.so
set -e
cd "$(dirname "$0")/../buas_pathtracer_b200/csrc"
[ -f build/host_scene.o ] || make -s >/dev/null
mkdir -p build/variants
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
$NVCC $ARCH -std=c++17 -O3 -lineinfo --fmad=false -Xcompiler -fPIC,-O2,-ffp-contract=off,-msse4.1,-fvisibility=hidden,-Wall $2 \
      -Xptxas -v -c bpt_device.cu -o build/variants/bpt_device_$1.o 2> build/variants/ptxas_$1.log
$NVCC $ARCH -shared -o build/variants/libbpt_$1.so build/host_scene.o build/bvh_build.o build/wide_bvh.o build/procedural_inputs.o build/obj_hdr_readers.o build/variants/bpt_device_$1.o -Xlinker --no-undefined -ldl
echo "built build/variants/libbpt_$1.so"
