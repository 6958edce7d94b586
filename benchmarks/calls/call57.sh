python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mel or mfcc or batch_forward_and_inverse or config1" 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-330
