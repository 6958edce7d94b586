timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "per_frame" 2>&1 | tail -3
