#!/bin/sh
# build_variant.sh OUT.so [nvcc flags...] -- an experimental build of libb200id.so next to the product one
# (used with B200ID_LIB=OUT.so for timing experiments; the product library is built by build.py)
set -e
here=$(cd "$(dirname "$0")" && pwd)
out=$1; shift
mkdir -p "$(dirname "$out")"
cd "$here/csrc"
nvcc -O3 -std=c++17 -shared -Xcompiler -fPIC -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -ffp-contract=off "$@" -o "$out" \
    capi.cu frf_project.cu gp_batch.cu gp_batch_tiled.cu gp_pinv.cu gp_large.cu chol_blocked.cu fir_conv.cu fir_fft.cu
