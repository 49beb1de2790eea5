#!/bin/bash
# Build libstif_b200.so in-tree for sm_100a (called by __graft_entry__.build()).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../lib"
mkdir -p "$OUT" "$HERE/obj"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
ARCH="-gencode arch=compute_100a,code=sm_100a"
CXXFLAGS="-O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-ffp-contract=off,-Wall"
pids=()
for f in stif_api.cu kernels_fp32.cu kernels_tc.cu kernels_hp.cu kernels_dcn.cu tc_selftest.cu; do
  ( "$NVCC" $ARCH $CXXFLAGS ${EXTRA_NVCC_FLAGS:-} -c "$HERE/$f" -o "$HERE/obj/${f%.cu}.o" ) &
  pids+=($!)
done
for f in pack_weights.cpp axis_tables.cpp host_plan.cpp; do
  ( "$NVCC" $ARCH $CXXFLAGS -c "$HERE/$f" -o "$HERE/obj/${f%.cpp}.o" ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
"$NVCC" $ARCH -shared -o "$OUT/libstif_b200.so" "$HERE"/obj/*.o
echo "built $OUT/libstif_b200.so"
