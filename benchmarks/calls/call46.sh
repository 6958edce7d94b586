timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "chunk or async or staged or stream_ordered or many_signals or pcm" 2>&1 | tail -2
timeout 600 python bench.py --steps 5 --warmup 3 --e2e-steps 6 --no-cpu-baseline --no-stream > gpurun_out/b46.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/b46.json')); print('e2e ms', round(d['e2e']['ms_per_step'],2), 'pcm16 ms', round(d['e2e_pcm16']['ms_per_step'],2))"
