set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "inverse or roundtrip or config or full" > gpurun_out/r02_gputests4.txt 2>&1; echo "rc=$?"
tail -3 gpurun_out/r02_gputests4.txt
timeout 300 python benchmarks/ab_kernels.py --rounds 10 --kinds inverse vv_dsp_b200/lib/libvvdsp_b200_r1.so vv_dsp_b200/lib/libvvdsp_b200_wspm0.so vv_dsp_b200/lib/libvvdsp_b200.so > gpurun_out/r02_ab4.jsonl 2> gpurun_out/r02_ab4.err; echo "rc=$?"
cat gpurun_out/r02_ab4.jsonl; tail -5 gpurun_out/r02_ab4.err
