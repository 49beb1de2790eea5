for d in 0 2000 4000 8000 10000 14000 18000; do STIF_DEPHASE_K1=$d STIF_DEPHASE_K2=$d timeout 120 python profiles/quick_bench.py; done
