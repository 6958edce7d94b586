mkdir -p gpurun_out
nvidia-smi -L | wc -l
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $T --nproc-per-node 8 --master-port 29571 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_final_n8.json 2> gpurun_out/r02_bench_final_n8.err; echo "bench n8 rc=$?"
timeout 600 $T --nproc-per-node 4 --master-port 29572 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r02_bench_final_n4.json 2> gpurun_out/r02_bench_final_n4.err; echo "bench n4 rc=$?"
for N in 8 4; do python - <<PY
import json
for l in open('gpurun_out/r02_bench_final_n$N.json'):
    if l.startswith('{'):
        d=json.loads(l); s=d.get('stream_config4') or {}
        print($N, 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'pcm16', round(d['e2e_pcm16']['value']), 'stream ms', s.get('ms_per_step'), 'eff', s.get('strong_scaling_efficiency'), 'bit', s.get('bit_identical_to_unsharded'), s.get('error'))
PY
done
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stream_sharding" 2>&1 | tail -2
