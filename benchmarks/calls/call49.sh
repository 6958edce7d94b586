timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mel" 2>&1 | tail -2
python benchmarks/logmel_e2e.py 2>/dev/null | cut -c1-300
VVB_STAGE_TARGET_BYTES=4000000000 python benchmarks/logmel_e2e.py 2>/dev/null | cut -c1-300
