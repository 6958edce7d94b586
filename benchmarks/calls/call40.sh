L=vv_dsp_b200/lib
python benchmarks/ab_kernels.py --rounds 8 --kinds inverse $L/libvvdsp_b200.so $L/libvvdsp_b200_r96.so $L/libvvdsp_b200_r112.so $L/libvvdsp_b200_direct.so | cut -c1-240
