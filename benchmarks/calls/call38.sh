mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests_final.txt 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r02_gputests_final.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench38.json 2> gpurun_out/r02_bench38.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02_bench38.json')); print(d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e_pcm16']['value'], d['stream_config4']['ms_per_step'])"
