set -x
mkdir -p gpurun_out
CMD="python benchmarks/ab_kernels.py --rounds 1 --reps 3 --warm 3 --kinds inverse vv_dsp_b200/lib/libvvdsp_b200.so"
$CMD > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:istft_ws -s 2 -c 1 -o gpurun_out/r02_ws_v1 $CMD > gpurun_out/ncu3.log 2>&1
tail -3 gpurun_out/ncu3.log
