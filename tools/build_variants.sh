#!/bin/bash
# Build experimental variants of libctxnerf.so (mlp_fwd / mlp_dgrad compiled with -D flags); select one at run
# time with CTXNERF_LIB=<path>.  Usage: tools/build_variants.sh name:"-DFLAG1 -DFLAG2" ...
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
PKG=$ROOT/contexture-nerf_b200
python $PKG/ctxnerf/build.py >/dev/null
mkdir -p $PKG/ctxnerf/variants
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  tmp=$(mktemp -d)
  for f in mlp_fwd mlp_dgrad; do
    nvcc -c $PKG/csrc/$f.cu -o $tmp/$f.o -I $ROOT/include -I $PKG/csrc -gencode arch=compute_100a,code=sm_100a \
      -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr $flags &
  done
  wait
  objs=$(ls $PKG/csrc/build/*.o | grep -v "mlp_fwd.o\|mlp_dgrad.o\|_diag.o")
  nvcc -shared -o $PKG/ctxnerf/variants/libctxnerf_$name.so $objs $tmp/mlp_fwd.o $tmp/mlp_dgrad.o -gencode arch=compute_100a,code=sm_100a
  rm -rf $tmp
  echo built $name
done
