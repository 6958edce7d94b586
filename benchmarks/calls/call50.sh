timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mel" 2>&1 | tail -2
python benchmarks/logmel_e2e.py 2>/dev/null | cut -c1-330 > gpurun_out/r02_logmel_e2e.jsonl; cat gpurun_out/r02_logmel_e2e.jsonl
