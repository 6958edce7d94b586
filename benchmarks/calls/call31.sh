L=vv_dsp_b200/lib
python benchmarks/ab_kernels.py --nfft 256 --hop 64 --rounds 5 --kinds inverse $L/libvvdsp_b200_old.so $L/libvvdsp_b200_swz.so $L/libvvdsp_b200.so | cut -c1-230
python benchmarks/ab_kernels.py --nfft 512 --hop 128 --rounds 5 --kinds inverse $L/libvvdsp_b200_old.so $L/libvvdsp_b200.so | cut -c1-230
