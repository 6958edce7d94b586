mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests9.txt 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r02_gputests9.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench9.json 2> gpurun_out/r02_bench9.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02_bench9.json')); print(json.dumps(d.get('stream_config4'), indent=1)); print(d['ms_per_step'], d['kernels'])"; tail -3 gpurun_out/r02_bench9.err
