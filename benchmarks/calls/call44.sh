mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-stream > gpurun_out/r02_bench44.json 2> gpurun_out/r02_bench44.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench44.json')); print(d['ms_per_step'], d['kernels']['stft_forward_ms'], d['kernels']['stft_inverse_ms'], d['roofline']['kernel'], d['roofline']['frac'], d['clocks'])"
python benchmarks/ab_kernels.py --rounds 6 --kinds complex,inverse vv_dsp_b200/lib/libvvdsp_b200.so | cut -c1-330
