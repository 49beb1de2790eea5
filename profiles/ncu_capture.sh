# round-1 ncu capture recipe.  One ncu pass per gpurun call, each only after the same command has exited 0 without ncu:
#   gpurun --timeout 600 -- 'bash profiles/ncu_capture.sh list'    -> gpurun_out/launches.csv   (launch list, per-launch times)
#   gpurun --timeout 600 -- 'bash profiles/ncu_capture.sh full'    -> gpurun_out/prof.ncu-rep   (--set full of one K0, K1, K2 launch)
# then, in the build container:  ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv; python profiles/ncu_raw_summary.py raw.csv
#                                ncu -i ... --page source --csv --kernel-name regex:k2_stage > src.csv; python profiles/ncu_src_summary.py src.csv
CMD="timeout 120 python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -3 gpurun_out/plain.log; exit 1; }
if [ "$1" = "full" ]; then
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:k[012]_ -s 10 -c 3 -o gpurun_out/prof $CMD > gpurun_out/ncu_full.log 2>&1
  tail -2 gpurun_out/ncu_full.log | cut -c1-160
else
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
  tail -2 gpurun_out/launches.csv | cut -c1-200
fi
