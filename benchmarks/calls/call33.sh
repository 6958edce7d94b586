mkdir -p gpurun_out
timeout 900 python benchmarks/sweep.py > gpurun_out/r02_sweep_config5_n1_final.jsonl 2> gpurun_out/sweep.err; echo "rc=$?"; wc -l gpurun_out/r02_sweep_config5_n1_final.jsonl; tail -2 gpurun_out/sweep.err
