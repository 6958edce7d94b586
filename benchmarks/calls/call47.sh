mkdir -p gpurun_out
L=vv_dsp_b200/lib/libvvdsp_b200.so
AB="python benchmarks/ab_kernels.py --rounds 1 --reps 3 --warm 3 $L"
$AB --kinds complex > gpurun_out/plain47.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_march -s 2 -c 1 -f -o gpurun_out/r02_final_fwd2048 $AB --kinds complex > gpurun_out/ncu47a.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:istft_ws -s 2 -c 1 -f -o gpurun_out/r02_final_inv2048 $AB --kinds inverse > gpurun_out/ncu47b.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_march -s 2 -c 1 -f -o gpurun_out/r02_final_logmel python benchmarks/logmel_bench.py > gpurun_out/ncu47c.log 2>&1; echo "rc=$?"
