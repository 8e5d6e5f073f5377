#!/bin/sh
# (dev) build an experimental variant of the library next to the product one:
#   sh tools/build_variant.sh NAME -DRD_EXP_FOO ...   ->  gpurun_ab/libramdsir_NAME.so   (use with RAMDSIR_LIB=...)
set -e
name="$1"; shift
root="$(cd "$(dirname "$0")/.." && pwd)"
mkdir -p "$root/gpurun_ab"
c="$root/ramdsir_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -t 4 -cudart shared -shared -Xcompiler -fPIC \
     -Xlinker -rpath=/usr/local/cuda/lib64 -o "$root/gpurun_ab/libramdsir_$name.so" \
     "$c/ramdsir_abi.cu" "$c/ram_fused.cu" "$c/ram_generic.cu" "$c/ram_multipass.cu" "$c/ram_select.cu" "$c/ram_transform.cu" "$@"
echo "built gpurun_ab/libramdsir_$name.so"
