python benchmarks/logmel_e2e.py 2>gpurun_out/logmel_e2e.err | cut -c1-420 > gpurun_out/r02_logmel_e2e.jsonl; cat gpurun_out/r02_logmel_e2e.jsonl; tail -3 gpurun_out/logmel_e2e.err
