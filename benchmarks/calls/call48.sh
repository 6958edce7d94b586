mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02_bench_final_v3.json 2> gpurun_out/b48.err; echo "rc=$?"; tail -2 gpurun_out/b48.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_final_v3.json')); print(d['ms_per_step'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['traffic_source'], d['e2e']['value'], d['e2e_pcm16']['value']); print(d['cpu_baseline']); print(d['cpu_baseline_1thread']); print(d['cpu_baseline_power']); print(d['clocks'])"
