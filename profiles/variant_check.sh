# Parity subset + timing of tuning variants:  bash profiles/variant_check.sh <variant>...   (first one also runs the parity subset)
mkdir -p gpurun_out
first=1
for v in "$@"; do
  if [ $first = 1 ]; then
    STIF_LIB=build_variants/$v.so timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "golden or config1 or config3 or properties or host_pipeline or fuzz or uint8 or band" 2>&1 | tail -3
    first=0
  fi
  STIF_LIB=build_variants/$v.so timeout 150 python profiles/quick_bench.py 2>&1 | tail -1 | sed "s/^/$v: /" | tee -a gpurun_out/variants.txt
done
