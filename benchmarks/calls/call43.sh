L=vv_dsp_b200/lib/libvvdsp_b200.so
python benchmarks/ab_kernels.py --hop 1024 --rounds 3 $L | cut -c1-330
python benchmarks/ab_kernels.py --hop 256 --batch 512 --rounds 3 $L | cut -c1-330
python benchmarks/ab_kernels.py --nfft 4096 --hop 2048 --rounds 3 $L | cut -c1-330
python benchmarks/ab_kernels.py --nfft 1024 --hop 512 --rounds 3 $L | cut -c1-330
