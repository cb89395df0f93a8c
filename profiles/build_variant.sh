#!/bin/bash
# Build a variant of libb200id.so into variants/<name>.so with extra nvcc flags (experiments: B200ID_LIB=variants/<name>.so).
#   profiles/build_variant.sh <name> [-DFLAG ...]
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$root/variants"
cd "$root/gauss-process_b200/csrc"
nvcc -O3 -std=c++17 -shared -Xcompiler -fPIC -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -ffp-contract=off \
     --threads 8 "$@" -o "$root/variants/$name.so" capi.cu frf_project.cu gp_batch.cu gp_batch_tiled.cu gp_pinv.cu gp_large.cu \
     chol_blocked.cu fir_conv.cu fir_fft.cu kr_fir.cu 2>&1 | grep -i -E "error|spill" || true
echo "built variants/$name.so"
