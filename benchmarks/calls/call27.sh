mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python benchmarks/logmel_bench.py 2>/dev/null | cut -c1-330
