mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "inverse or roundtrip or config or full" 2>&1 | tail -2
python benchmarks/ab_kernels.py --rounds 10 --kinds inverse vv_dsp_b200/lib/libvvdsp_b200_r1.so vv_dsp_b200/lib/libvvdsp_b200_wsv2.so vv_dsp_b200/lib/libvvdsp_b200.so
echo burst; python benchmarks/ab_kernels.py --rounds 1 --reps 3 --warm 3 --kinds inverse vv_dsp_b200/lib/libvvdsp_b200.so
