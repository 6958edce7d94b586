export LOGMEL_SMALL_FUSED_ONLY=1
LOGMEL_SMALL_NFFT=400 python benchmarks/logmel_small.py > gpurun_out/plain53.log 2>&1; echo "plain rc=$?"; cat gpurun_out/plain53.log | tail -2
LOGMEL_SMALL_NFFT=400 timeout 300 ncu --set full --clock-control none --import-source on -k regex:stft_forward_kernel -s 3 -c 1 -f -o gpurun_out/r02_logmel_generic_400 python benchmarks/logmel_small.py > gpurun_out/ncu53a.log 2>&1; echo "rc=$?"
LOGMEL_SMALL_NFFT=1024 timeout 300 ncu --set full --clock-control none --import-source on -k regex:stft_forward_kernel -s 3 -c 1 -f -o gpurun_out/r02_logmel_generic_1024 python benchmarks/logmel_small.py > gpurun_out/ncu53b.log 2>&1; echo "rc=$?"
