mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests10.txt 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02_gputests10.txt
for sz in "4096 1024" "8192 2048"; do set -- $sz
python benchmarks/ab_kernels.py --nfft $1 --hop $2 --rounds 6 vv_dsp_b200/lib/libvvdsp_b200_r1.so vv_dsp_b200/lib/libvvdsp_b200_tw3off.so vv_dsp_b200/lib/libvvdsp_b200.so
done
