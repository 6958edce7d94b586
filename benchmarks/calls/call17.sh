mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mel or mfcc or bluestein_above" 2>&1 | tail -5
python benchmarks/logmel_bench.py 2>/dev/null | cut -c1-420
VVB_MEL_SINGLE=1 python benchmarks/logmel_bench.py 2>/dev/null | head -1 | cut -c1-420
CMD="python benchmarks/logmel_bench.py"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_march -s 2 -c 1 -f -o gpurun_out/r02_fused_logmel_v2 $CMD > gpurun_out/ncu17.log 2>&1; echo "ncu rc=$?"
