timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fallback or fused_in_generic or mel_host_pipeline" 2>&1 | tail -3
