timeout 120 python profiles/quick_bench.py
timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
rm -f gpurun_out/trace.txt
STIF_TRACE=gpurun_out/trace.txt timeout 300 python profiles/trace_run.py
