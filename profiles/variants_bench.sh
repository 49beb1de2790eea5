# Time tuning variants (build_variants/<name>.so, profiles/build_variant.sh) at config 2 and check each against the default build's output:
#   gpurun --timeout 900 -- 'bash profiles/variants_bench.sh A B C ...'   -> gpurun_out/variants.txt
mkdir -p gpurun_out
for v in "$@"; do
  STIF_LIB=build_variants/$v.so timeout 150 python profiles/quick_bench.py 2>&1 | tail -1 | sed "s/^/$v: /" | tee -a gpurun_out/variants.txt
done
