mkdir -p gpurun_out
nvidia-smi -L | wc -l
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $T --nproc-per-node 2 --master-port 29561 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench24_n2.json 2> gpurun_out/r02_bench24_n2.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/r02_bench24_n2.err
python - <<'PY'
import json
for l in open('gpurun_out/r02_bench24_n2.json'):
    if l.startswith('{'):
        d=json.loads(l); s=d.get('stream_config4') or {}
        print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'pcm16', round(d['e2e_pcm16']['value']), 'stream', s.get('ms_per_step'), s.get('strong_scaling_efficiency'), s.get('bit_identical_to_unsharded'), s.get('error'))
PY
timeout 900 $T --nproc-per-node 2 --master-port 29562 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stream_sharding" 2>&1 | tail -3
