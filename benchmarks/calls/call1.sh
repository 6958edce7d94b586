set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
./benchmarks/micro/microbench > gpurun_out/r02_microbench.jsonl 2>&1
python benchmarks/ab_kernels.py --rounds 5 vv_dsp_b200/lib/libvvdsp_b200_r1.so vv_dsp_b200/lib/libvvdsp_b200_pm0.so vv_dsp_b200/lib/libvvdsp_b200.so vv_dsp_b200/lib/libvvdsp_b200_pm2.so > gpurun_out/r02_ab1.jsonl 2> gpurun_out/r02_ab1.err
python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests1.txt 2>&1
tail -3 gpurun_out/r02_gputests1.txt
cat gpurun_out/r02_ab1.jsonl
