export LOGMEL_SMALL_NFFT=400
for i in 1 2; do
python benchmarks/logmel_small.py 2>>gpurun_out/ab58.err | cut -c60-260
VVDSP_B200_LIB=$PWD/vv_dsp_b200/lib/ab/libvvdsp_b200_nf8.so python benchmarks/logmel_small.py 2>>gpurun_out/ab58.err | cut -c60-260
done
tail -3 gpurun_out/ab58.err
