python benchmarks/logmel_small.py 2>gpurun_out/logmel_small.err > gpurun_out/r02_logmel_small.jsonl; cut -c1-330 gpurun_out/r02_logmel_small.jsonl; tail -3 gpurun_out/logmel_small.err
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02_gputests_final.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r02_bench_final_v4.json 2>gpurun_out/bench.err; cut -c1-600 gpurun_out/r02_bench_final_v4.json
