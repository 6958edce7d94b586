mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python benchmarks/logmel_bench.py > gpurun_out/r02_logmel_mfcc.jsonl 2>gpurun_out/logmel.err; cat gpurun_out/r02_logmel_mfcc.jsonl; tail -2 gpurun_out/logmel.err
VVB_MEL_NO_SCAN=1 python benchmarks/logmel_bench.py 2>/dev/null | head -1 | cut -c1-300
python benchmarks/ab_kernels.py --rounds 8 --kinds complex,power vv_dsp_b200/lib/libvvdsp_b200_r1.so vv_dsp_b200/lib/libvvdsp_b200.so
python benchmarks/ab_kernels.py --nfft 4096 --hop 1024 --rounds 5 --kinds power vv_dsp_b200/lib/libvvdsp_b200_r1.so vv_dsp_b200/lib/libvvdsp_b200.so
