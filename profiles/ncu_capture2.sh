# round-2 ncu capture: plain run, launch list, then --set full of one launch of each kernel -> gpurun_out/<tag>_*
TAG=${1:-r02}
CMD="timeout 120 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sharded"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -3 gpurun_out/${TAG}_plain.log; exit 1; }
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_list.log 2>&1
tail -3 gpurun_out/${TAG}_launches.csv | cut -c1-200
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k[012]_ -s 10 -c 3 -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log | cut -c1-160
