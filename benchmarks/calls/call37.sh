mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "warp_specialised_4096" 2>&1 | tail -3
L=vv_dsp_b200/lib/libvvdsp_b200.so
python benchmarks/ab_kernels.py --nfft 4096 --hop 1024 --rounds 6 --kinds inverse $L $L@VVB_WS3=1 | cut -c1-250
python benchmarks/ab_kernels.py --nfft 4096 --hop 512 --batch 512 --rounds 3 --kinds inverse $L $L@VVB_WS3=1 | cut -c1-250
CMD="python benchmarks/ab_kernels.py --nfft 4096 --hop 1024 --rounds 1 --reps 3 --warm 3 --kinds inverse $L@VVB_WS3=1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:istft_ws3 -s 2 -c 1 -f -o gpurun_out/r02_inv4096_ws3 $CMD > gpurun_out/ncu37.log 2>&1; echo "ncu rc=$?"
