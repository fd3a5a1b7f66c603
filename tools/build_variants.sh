#!/bin/bash
# usage: tools/build_variants.sh "name1:-DFOO=1 -DBAR=2" "name2:..."   -> build/variants/<name>.so
# (the -D switches reach the two raster translation units, oit_raster.cu and oit_raster_ll.cu; everything else is shared)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
CS=$ROOT/vk_order_independent_transparency_b200/csrc
make -C $CS -j8 -s
mkdir -p $ROOT/build/variants
rm -f $ROOT/build/variants/*.so
build_one() {
  name=${1%%:*}; flags=${1#*:}
  FL="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off"
  nvcc $FL $flags -c $CS/oit_raster_ll.cu -o /tmp/variant_${name}_ll.o
  if echo "$flags" | grep -q "OIT_LL_"; then cp $CS/oit_raster.o /tmp/variant_$name.o; else nvcc $FL $flags -c $CS/oit_raster.cu -o /tmp/variant_$name.o; fi
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $ROOT/build/variants/$name.so /tmp/variant_$name.o /tmp/variant_${name}_ll.o $CS/oit_raster_q.o $CS/oit_api.o $CS/oit_geometry.o $CS/oit_composite.o $CS/oit_gather.o $CS/oit_peer.o $CS/oit_scene.o -lcudart -ldl
  echo built $name
}
export -f build_one; export ROOT CS
printf '%s\n' "$@" | xargs -P 8 -I{} bash -c 'build_one "{}"'
