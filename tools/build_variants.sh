#!/bin/bash
# Build variants of librb200.so into reina-vk_b200/csrc/variants/ (git-ignored; they travel to the GPU box with gpurun).
# usage: tools/build_variants.sh name1="-DFLAG=1 -DOTHER=2" name2="..."
cd "$(dirname "$0")/../reina-vk_b200/csrc"
mkdir -p variants
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xptxas -v -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -Xcompiler -ffp-contract=off"
build_one() {
  name="${1%%=*}"; extra="${1#*=}"
  d=$(mktemp -d)
  for f in api bvh_build wavefront post group; do
    nvcc $FLAGS $extra -c $f.cu -o $d/$f.o 2> variants/$name.$f.log || { echo "BUILD FAILED: $name ($f)"; tail -5 variants/$name.$f.log; rm -rf $d; return 1; }
  done
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/$name.so $d/api.o $d/bvh_build.o $d/wavefront.o $d/post.o $d/group.o -ldl && echo "built variants/$name.so [$extra]"
  grep -A2 "k_extendILb0\|k_shadowILb0" variants/$name.wavefront.log | grep -E "spill|Used" | tr '\n' ' '; echo
  rm -rf $d
}
for v in "$@"; do build_one "$v" & done
wait
