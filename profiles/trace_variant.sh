# clock64 timeline of a variant built with -DSTIF_ENABLE_TRACE:  bash profiles/trace_variant.sh <variant> -> gpurun_out/trace_<variant>.txt
for v in "$@"; do
rm -f gpurun_out/trace_$v.txt
STIF_LIB=build_variants/$v.so STIF_TRACE=gpurun_out/trace_$v.txt timeout 300 python profiles/trace_run.py
STIF_LIB=build_variants/$v.so timeout 150 python profiles/quick_bench.py 2>&1 | tail -1 | sed "s/^/$v (trace build): /"
done
