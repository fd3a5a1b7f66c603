#!/bin/bash
# usage: tools/build_variants.sh "name1:-DFOO=1 -DBAR=2" "name2:..."   -> build/variants/<name>.so (only oit_raster.cu differs)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
CS=$ROOT/vk_order_independent_transparency_b200/csrc
make -C $CS -j8 -s
mkdir -p $ROOT/build/variants
rm -f $ROOT/build/variants/*.so
build_one() {
  name=${1%%:*}; flags=${1#*:}
  # a non-default CTA size changes oit_internal.h constants used by every TU? only oit_raster.cu uses RASTER_THREADS
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off $flags \
       -c $CS/oit_raster.cu -o /tmp/variant_$name.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $ROOT/build/variants/$name.so /tmp/variant_$name.o $CS/oit_api.o $CS/oit_geometry.o $CS/oit_composite.o $CS/oit_gather.o $CS/oit_peer.o $CS/oit_scene.o -lcudart -ldl
  echo built $name
}
export -f build_one; export ROOT CS
printf '%s\n' "$@" | xargs -P 8 -I{} bash -c 'build_one "{}"'
