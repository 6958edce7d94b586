mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests_final.txt 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r02_gputests_final.txt
timeout 900 python bench.py --impl reference > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/r02_bench_reference_arm.json
timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_bench_final.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_final.json'))
print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), d['kernels'])
print('roofline', d['roofline'])
print('e2e', round(d['e2e']['value']), 'pcm16', round(d['e2e_pcm16']['value']), 'cpu', d['cpu_baseline']['value'], d['cpu_baseline_1thread']['value'], 'launches', d['gpu_launches'], d['clocks'])
print('stream', d.get('stream_config4'))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_lf.log 2>&1; echo "launch list rc=$?"
