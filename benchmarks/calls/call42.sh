mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "batch_forward_and_inverse or inverse_few or marching_istft or golden or full_size or stream_sharding or many_signals" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench42.json 2> gpurun_out/r02_bench42.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02_bench42.json')); print(d['ms_per_step'], d['kernels'], d['roofline']['frac'], d['clocks'])"
