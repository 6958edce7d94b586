mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pcm or random_filterbanks" 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stream > gpurun_out/r02_bench22.json 2> gpurun_out/r02_bench22.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench22.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench22.json')); print(d['ms_per_step'], d['e2e'], d['e2e_pcm16'])"
export PYTHONUNBUFFERED=1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mel_fused or test_mel or marching_istft or short_and_ragged or batch_forward_and_inverse or pcm" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/sanitizer_memcheck.log; tail -4 gpurun_out/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mel_fused_kernel" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -c "hazard" gpurun_out/sanitizer_racecheck.log; grep "hazard" gpurun_out/sanitizer_racecheck.log | cut -c1-200 | sort | uniq -c | sort -rn | head -8; tail -4 gpurun_out/sanitizer_racecheck.log
