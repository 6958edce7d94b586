mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_c_callers.py -m gpu -x -q -k "fft or per_frame or known_answers or c_callers or bluestein_sizes or mixed" 2>&1 | tail -3
python benchmarks/perframe_latency.py > gpurun_out/r02_perframe_latency.jsonl 2>/dev/null; VVB_PERFRAME_STAGED=1 python benchmarks/perframe_latency.py 2>/dev/null | head -6 | sed 's/per-frame API)/per-frame API, VVB_PERFRAME_STAGED=1: copy + launch + copy)/' >> gpurun_out/r02_perframe_latency.jsonl; cut -c1-150 gpurun_out/r02_perframe_latency.jsonl
