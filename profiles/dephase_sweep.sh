for d in 0 6000 12000 -12000; do STIF_DEPHASE_K1=$d STIF_DEPHASE_K2=$d timeout 120 python profiles/quick_bench.py; done
rm -f gpurun_out/trace.txt
STIF_DEPHASE_K1=14000 STIF_DEPHASE_K2=12000 STIF_TRACE=gpurun_out/trace.txt timeout 300 python profiles/trace_run.py
