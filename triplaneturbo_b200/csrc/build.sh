#!/usr/bin/env bash
# Builds libtriplane_b200.so in-tree for sm_100a.  Usage: [TT_LIBNAME=name.so] build.sh [extra nvcc flags]
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
out="$here/../lib"
mkdir -p "$out"
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo \
     -Xcompiler -fPIC -shared -Xptxas -v "$@" \
     -o "$out/${TT_LIBNAME:-libtriplane_b200.so}" "$here/tt_kernels.cu" 2> "$out/ptxas.log" || { cat "$out/ptxas.log" >&2; exit 1; }
grep -E "error|warning" "$out/ptxas.log" | grep -v "Wno" >&2 || true
