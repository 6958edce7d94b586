python benchmarks/logmel_small.py 2>gpurun_out/logmel_small.err > gpurun_out/r02_logmel_small.jsonl; cut -c1-420 gpurun_out/r02_logmel_small.jsonl; tail -3 gpurun_out/logmel_small.err
