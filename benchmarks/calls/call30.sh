mkdir -p gpurun_out
L=vv_dsp_b200/lib/libvvdsp_b200.so
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "batch_forward_and_inverse or inverse_few or marching_istft or golden or voicebank or short_and_ragged" 2>&1 | tail -2
python benchmarks/ab_kernels.py --nfft 256 --hop 64 --rounds 4 --kinds inverse $L | cut -c1-260
python benchmarks/ab_kernels.py --nfft 512 --hop 128 --rounds 4 --kinds inverse $L | cut -c1-260
python benchmarks/ab_kernels.py --nfft 256 --hop 128 --rounds 3 --kinds inverse $L | cut -c1-260
python benchmarks/ab_kernels.py --nfft 512 --hop 64 --rounds 3 --kinds inverse $L | cut -c1-260
