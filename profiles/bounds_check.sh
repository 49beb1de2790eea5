# The parity suite against a build with device-side bounds asserts (-DSTIF_CHECK_BOUNDS): substitute for compute-sanitizer, which is closed on this pool.
STIF_LIB=build_variants/BC.so timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/bounds_check.log
