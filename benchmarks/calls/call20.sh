mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "full_size" 2>&1 | tail -3
python benchmarks/nonpow2_bench.py > gpurun_out/r02_nonpow2.jsonl 2>gpurun_out/nonpow2.err; cut -c1-330 gpurun_out/r02_nonpow2.jsonl; tail -2 gpurun_out/nonpow2.err
python benchmarks/perframe_latency.py > gpurun_out/r02_perframe_latency.jsonl 2>gpurun_out/perframe.err; cut -c1-300 gpurun_out/r02_perframe_latency.jsonl; tail -2 gpurun_out/perframe.err
python benchmarks/logmel_bench.py > gpurun_out/r02_logmel_mfcc.jsonl 2>/dev/null; VVB_MEL_UNFUSED=1 python benchmarks/logmel_bench.py >> gpurun_out/r02_logmel_mfcc.jsonl 2>/dev/null; cut -c1-200 gpurun_out/r02_logmel_mfcc.jsonl
