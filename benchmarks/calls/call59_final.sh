python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r02_smoke_final.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02_gputests_final.txt
python benchmarks/logmel_small.py 2>gpurun_out/logmel_small.err > gpurun_out/r02_logmel_small.jsonl; cut -c1-260 gpurun_out/r02_logmel_small.jsonl
