#!/bin/bash
# A/B builds: compile the CUDA library of another git revision into wavenets_b200/libwavenet_b200_<tag>.so (load it with WN_LIB=...)
# usage: scripts/build_rev.sh <rev> <tag> [extra nvcc flags]
set -e
REV=$1; TAG=$2; shift 2
D=$(mktemp -d)
git archive $REV wavenets_b200/csrc include | tar -x -C $D
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared --expt-relaxed-constexpr -ldl "$@" \
  $D/wavenets_b200/csrc/wn_api.cu -o wavenets_b200/libwavenet_b200_$TAG.so
rm -rf $D
echo wavenets_b200/libwavenet_b200_$TAG.so
