mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
# 1. concurrent host-link probe (the e2e floor)
python benchmarks/pcie_probe.py > gpurun_out/r02_pcie_probe.jsonl 2>gpurun_out/pcie.err
for N in 2 4 8; do $T --nproc-per-node $N --master-port 2951$N benchmarks/pcie_probe.py >> gpurun_out/r02_pcie_probe.jsonl 2>>gpurun_out/pcie.err; done
$T --nproc-per-node 8 --master-port 29519 benchmarks/pcie_probe.py --affinity >> gpurun_out/r02_pcie_probe.jsonl 2>>gpurun_out/pcie.err
cat gpurun_out/r02_pcie_probe.jsonl
# 2. single stream, C handle, 1/2/4/8 GPUs (graphs and plain enqueue)
timeout 600 python benchmarks/stream_bench.py > gpurun_out/r02_stream_config4_scaling.jsonl 2> gpurun_out/stream.err; cat gpurun_out/r02_stream_config4_scaling.jsonl | cut -c 150-700; tail -2 gpurun_out/stream.err
timeout 600 python benchmarks/stream_bench.py --no-graph --gpus 1 8 > gpurun_out/r02_stream_config4_nograph.jsonl 2>> gpurun_out/stream.err; cat gpurun_out/r02_stream_config4_nograph.jsonl | cut -c 150-400
# 3. multi-process flavour over NCCL, 8 ranks
timeout 300 $T --nproc-per-node 8 --master-port 29520 tests/nccl_stream_check.py > gpurun_out/r02_nccl_stream_check.txt 2>&1; tail -3 gpurun_out/r02_nccl_stream_check.txt
# 4. headline bench at 1/2/4/8 (weak scaling) like the driver's SCALE run
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_scale_n1.json 2> gpurun_out/scale.err
for N in 2 4 8; do timeout 600 $T --nproc-per-node $N --master-port 2953$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_scale_n$N.json 2>> gpurun_out/scale.err; done
for N in 1 2 4 8; do python - <<PY
import json
for l in open('gpurun_out/r02_scale_n$N.json'):
    if l.startswith('{'):
        d=json.loads(l); s=d.get('stream_config4') or {}
        print($N, 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'stream ms', s.get('ms_per_step'), 'eff', s.get('strong_scaling_efficiency'), 'bit', s.get('bit_identical_to_unsharded'), s.get('error'))
PY
done
# 5. config 5 sweep at 1/2/4/8
timeout 900 python benchmarks/sweep.py > gpurun_out/r02_sweep_config5_n1.jsonl 2> gpurun_out/sweep.err
for N in 2 4 8; do timeout 900 $T --nproc-per-node $N --master-port 2954$N benchmarks/sweep.py > gpurun_out/r02_sweep_config5_n$N.jsonl 2>> gpurun_out/sweep.err; done
wc -l gpurun_out/r02_sweep_config5_n*.jsonl; tail -3 gpurun_out/sweep.err
