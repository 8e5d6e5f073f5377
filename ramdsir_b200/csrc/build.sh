#!/bin/sh
# Build libramdsir.so in-tree for sm_100a (the only target).  Usage: sh ramdsir_b200/csrc/build.sh
set -e
here="$(cd "$(dirname "$0")" && pwd)"
out="$here/../libramdsir.so"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -t 4 -cudart shared \
     -shared -Xcompiler -fPIC -Xlinker -rpath=/usr/local/cuda/lib64 \
     -o "$out" "$here/ramdsir_abi.cu" "$here/ram_fused.cu" "$here/ram_generic.cu" "$here/ram_multipass.cu" "$here/ram_select.cu" "$here/ram_transform.cu" "$@"
echo "built $out"
