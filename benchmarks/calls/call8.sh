mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests8.txt 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r02_gputests8.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench8.json 2> gpurun_out/r02_bench8.err; echo "bench rc=$?"; cat gpurun_out/r02_bench8.json; tail -3 gpurun_out/r02_bench8.err
