#!/bin/bash
# Build a tuning variant of the library:  bash profiles/build_variant.sh <name> "<nvcc -D flags>"  -> build_variants/<name>.so
# (only kernels_tc.cu is recompiled; the other objects come from the last full build.)  Select it with STIF_LIB=build_variants/<name>.so.
set -euo pipefail
NAME=$1; FLAGS=${2:-}
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
C="$ROOT/stif-continuous-video-representation_b200/csrc"
mkdir -p "$ROOT/build_variants/obj_$NAME"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
ARCH="-gencode arch=compute_100a,code=sm_100a"
"$NVCC" $ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-ffp-contract=off,-Wall $FLAGS -c "$C/kernels_tc.cu" -o "$ROOT/build_variants/obj_$NAME/kernels_tc.o"
OBJS=$(ls "$C"/obj/*.o | grep -v kernels_tc.o)
"$NVCC" $ARCH -shared -o "$ROOT/build_variants/$NAME.so" $OBJS "$ROOT/build_variants/obj_$NAME/kernels_tc.o"
echo "built build_variants/$NAME.so ($FLAGS)"
