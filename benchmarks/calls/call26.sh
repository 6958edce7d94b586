mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mixed_radix or bluestein" 2>&1 | tail -3
python benchmarks/nonpow2_bench.py > gpurun_out/r02_nonpow2.jsonl 2>/dev/null; cut -c1-330 gpurun_out/r02_nonpow2.jsonl
