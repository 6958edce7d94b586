timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mel or mfcc" 2>&1 | tail -3
python benchmarks/logmel_small.py 2>gpurun_out/logmel_small.err > gpurun_out/r02_logmel_small.jsonl; cut -c1-420 gpurun_out/r02_logmel_small.jsonl; tail -3 gpurun_out/logmel_small.err
