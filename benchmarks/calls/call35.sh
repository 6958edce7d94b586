mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_c_callers.py -m gpu -x -q -k "per_frame or known_answers or c_callers or status or reconstruct or spectrogram" 2>&1 | tail -3
python benchmarks/perframe_latency.py 2>/dev/null | head -6 | cut -c1-200
VVB_PERFRAME_STAGED=1 python benchmarks/perframe_latency.py 2>/dev/null | head -6 | cut -c1-200
