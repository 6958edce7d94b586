mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "warp_specialised" 2>&1 | tail -3
L=vv_dsp_b200/lib/libvvdsp_b200.so
python benchmarks/ab_kernels.py --rounds 6 --kinds complex,power $L $L@VVB_FWD_WS=1
python benchmarks/ab_kernels.py --rounds 4 --hop 256 --batch 512 --kinds complex,power $L $L@VVB_FWD_WS=1
CMD="python benchmarks/ab_kernels.py --rounds 1 --reps 3 --warm 3 --kinds complex $L@VVB_FWD_WS=1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_ws -s 2 -c 1 -f -o gpurun_out/r02_fwd_ws_v1 $CMD > gpurun_out/ncu23.log 2>&1; echo "ncu rc=$?"
