mkdir -p gpurun_out
L=vv_dsp_b200/lib/libvvdsp_b200.so
AB="python benchmarks/ab_kernels.py --rounds 1 --reps 3 --warm 3 $L"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:istft_pair -s 2 -c 1 -f -o gpurun_out/r02_inv256 $AB --nfft 256 --hop 64 --kinds inverse > gpurun_out/ncu29a.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:istft_pair -s 2 -c 1 -f -o gpurun_out/r02_inv512 $AB --nfft 512 --hop 128 --kinds inverse > gpurun_out/ncu29b.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_forward -s 2 -c 1 -f -o gpurun_out/r02_fwd256p $AB --nfft 256 --hop 64 --kinds power > gpurun_out/ncu29c.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_forward -s 2 -c 1 -f -o gpurun_out/r02_fwd1024 $AB --nfft 1024 --hop 256 --kinds complex > gpurun_out/ncu29d.log 2>&1; echo "rc=$?"
