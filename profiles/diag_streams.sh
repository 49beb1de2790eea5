# What each memory stream of K1 / K2 costs end to end: the library rebuilt with one stream compiled out (STIF_DIAG bit 0: Q-table
# stores, 1: TA loads, 2: stage-B TB loads, 3: K2 Q-table tap loads, 4: K2 TE tap loads; results are wrong by construction), kernels timed with quick_bench.py.
#   in the build container:  for d in 1 2 4 8; do EXTRA_NVCC_FLAGS=-DSTIF_DIAG=$d bash .../csrc/build.sh; cp .../lib/libstif_b200.so build_variants/diag$d.so; done
#   on the GPU box:          bash profiles/diag_streams.sh
L=stif-continuous-video-representation_b200/lib/libstif_b200.so
cp $L /tmp/libstif_keep.so
for d in ${DIAGS:-0 1 2 4 8 16 24 0}; do
  [ $d = 0 ] && cp /tmp/libstif_keep.so $L || cp build_variants/diag$d.so $L
  timeout 100 python profiles/quick_bench.py 2>&1 | tail -1 | sed "s/^/STIF_DIAG=$d /"
done
cp /tmp/libstif_keep.so $L
