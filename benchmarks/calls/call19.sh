mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mel or mfcc" 2>&1 | tail -2
python benchmarks/logmel_bench.py 2>/dev/null | cut -c1-300
CMD="python benchmarks/logmel_bench.py"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_march -s 2 -c 1 -f -o gpurun_out/r02_fused_logmel_v5 $CMD > gpurun_out/ncu19.log 2>&1; echo "ncu rc=$?"
