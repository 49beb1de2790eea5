# rebuild with the clock64 trace marks compiled in (GPU-box copy only), then dump the timeline of CTA 0
EXTRA_NVCC_FLAGS=-DSTIF_ENABLE_TRACE ./stif-continuous-video-representation_b200/csrc/build.sh > /dev/null 2>&1
rm -f gpurun_out/trace.txt
STIF_TRACE=gpurun_out/trace.txt timeout 300 python profiles/trace_run.py
