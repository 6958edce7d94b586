mkdir -p gpurun_out
python benchmarks/ab_kernels.py --nfft 4096 --hop 1024 --rounds 6 vv_dsp_b200/lib/libvvdsp_b200_r1.so vv_dsp_b200/lib/libvvdsp_b200_nohx.so vv_dsp_b200/lib/libvvdsp_b200.so
python benchmarks/stream_bench.py --gpus 1 | cut -c150-420
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
