mkdir -p gpurun_out
L=vv_dsp_b200/lib/libvvdsp_b200.so
echo "burst alone";     python benchmarks/ab_kernels.py --rounds 1 --reps 3 --warm 3 --kinds inverse $L
echo "sustained alone"; python benchmarks/ab_kernels.py --rounds 10 --kinds inverse $L
echo "sustained A/B";   python benchmarks/ab_kernels.py --rounds 10 --kinds inverse vv_dsp_b200/lib/libvvdsp_b200_r1.so $L
echo "sustained alone long"; python benchmarks/ab_kernels.py --rounds 40 --kinds inverse,complex $L
