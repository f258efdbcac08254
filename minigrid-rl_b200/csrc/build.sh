#!/bin/sh
# Builds the C-ABI CUDA library in-tree for sm_100a (cross-compiles without a GPU).  The translation units are
# compiled in parallel, then linked.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
OUT="$HERE/../_lib"
OBJ="$OUT/obj"
mkdir -p "$OUT" "$OBJ"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -I $ROOT/include -I $HERE"
pids=""
for f in mgrl_kernels mgrl_policy mgrl_policy_tc mgrl_update mgrl_wire mgrl_linear_tc5 mgrl_conv1_tc5 $MGRL_EXTRA_UNITS; do
    [ -f "$HERE/$f.cu" ] || continue
    "$NVCC" $FLAGS "$@" -c -o "$OBJ/$f.o" "$HERE/$f.cu" &
    pids="$pids $!"
done
# host half of the PCIe wire format (SSSE3 byte shuffles; picked at run time only when the CPU has them)
${CXX:-g++} -O3 -std=c++17 -mssse3 -fPIC -c -o "$OBJ/mgrl_wire_host.o" "$HERE/mgrl_wire_host.cpp" &
pids="$pids $!"
for p in $pids; do wait "$p"; done
"$NVCC" -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT/libmgrl.so" "$OBJ"/*.o
echo "built $OUT/libmgrl.so"
