mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/r02_bench15.json 2> gpurun_out/r02_bench15.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_bench15.err
# launch list of the same command shape (short run)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches15.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_l15.log 2>&1; echo "launch list rc=$?"
AB="python benchmarks/ab_kernels.py --rounds 1 --reps 3 --warm 3 vv_dsp_b200/lib/libvvdsp_b200.so"
$AB --kinds complex > gpurun_out/plain15.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_march -s 2 -c 1 -f -o gpurun_out/r02_fwd2048 $AB --kinds complex > gpurun_out/ncu15a.log 2>&1; echo "ncu fwd rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:istft_ws -s 2 -c 1 -f -o gpurun_out/r02_inv2048 $AB --kinds inverse > gpurun_out/ncu15b.log 2>&1; echo "ncu inv rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_march -s 2 -c 1 -f -o gpurun_out/r02_fwd4096 $AB --nfft 4096 --hop 1024 --kinds complex > gpurun_out/ncu15c.log 2>&1; echo "ncu fwd4096 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:istft -s 2 -c 1 -f -o gpurun_out/r02_inv4096 $AB --nfft 4096 --hop 1024 --kinds inverse > gpurun_out/ncu15d.log 2>&1; echo "ncu inv4096 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_march -s 2 -c 1 -f -o gpurun_out/r02_fwd8192 $AB --nfft 8192 --hop 2048 --kinds complex > gpurun_out/ncu15e.log 2>&1; echo "ncu fwd8192 rc=$?"
python benchmarks/logmel_bench.py 2>/dev/null | cut -c1-260
