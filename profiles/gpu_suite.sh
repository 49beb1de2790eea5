# Full GPU check used at milestones:  gpurun --timeout 1500 -- 'bash profiles/gpu_suite.sh <tag>'
# -> gpurun_out/<tag>_pytest.log, <tag>_bench.json, <tag>_bench_ref.json
TAG=${1:-run}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > gpurun_out/${TAG}_smi.txt 2>&1
nproc >> gpurun_out/${TAG}_smi.txt; free -g | head -2 >> gpurun_out/${TAG}_smi.txt
timeout 1300 python -m pytest tests -m gpu -q -x -s --durations=15 > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; cut -c1-600 gpurun_out/${TAG}_bench.json
