#!/usr/bin/env bash
# TEST TOOLING: compiles the CUDA kernel sources for the host with the emulation header (see cuda_emul.h).
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
src="$here/../../triplaneturbo_b200/csrc/tt_kernels.cu"
g++ -O2 -std=c++20 -ffp-contract=off -fPIC -shared -pthread -DTT_EMUL ${TT_EMUL_FLAGS:-} -include "$here/cuda_emul.h" \
    -x c++ "$src" -o "$here/libtt_emul.so"
