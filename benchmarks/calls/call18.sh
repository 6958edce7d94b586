mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python benchmarks/logmel_bench.py 2>/dev/null | cut -c1-420
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench18.json 2> gpurun_out/r02_bench18.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02_bench18.json')); print(d['ms_per_step'], d['kernels'], d['roofline']['frac'], d['roofline']['traffic'])"
