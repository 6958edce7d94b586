mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mixed_radix or bluestein or batch_forward_and_inverse or per_frame or fft_plans or fft_execute" 2>&1 | tail -3
python benchmarks/nonpow2_bench.py 2>/dev/null | head -2 | cut -c1-330
VVB_NO_MIXED_RADIX=1 python benchmarks/nonpow2_bench.py 2>/dev/null | head -1 | cut -c1-330
python __graft_entry__.py smoke 2>&1 | tail -2
