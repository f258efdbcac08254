#!/bin/sh
# Builds the C-ABI CUDA library in-tree for sm_100a (cross-compiles without a GPU).
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
OUT="$HERE/../_lib"
mkdir -p "$OUT"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
"$NVCC" -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo \
    -Xcompiler -fPIC -shared -I "$ROOT/include" -I "$HERE" \
    -o "$OUT/libmgrl.so" "$HERE/mgrl_kernels.cu" "$HERE/mgrl_policy.cu" "$HERE/mgrl_policy_tc.cu" "$@"
echo "built $OUT/libmgrl.so"
