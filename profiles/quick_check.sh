timeout 120 python profiles/quick_bench.py
timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
