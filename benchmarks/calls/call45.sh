mkdir -p gpurun_out
for sz in 201326592 100663296 50331648 25165824; do
VVB_STAGE_TARGET_BYTES=$sz timeout 600 python bench.py --steps 3 --warmup 3 --e2e-steps 6 --no-cpu-baseline --no-stream > gpurun_out/b45.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/b45.json')); print($sz, 'e2e ms', round(d['e2e']['ms_per_step'],2), 'pcm16 ms', round(d['e2e_pcm16']['ms_per_step'],2))"
done
