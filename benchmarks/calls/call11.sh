mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests11.txt 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02_gputests11.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench11_n2.json 2> gpurun_out/r02_bench11_n2.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02_bench11_n2.json')); print(json.dumps(d.get('stream_config4'), indent=1)); print(d['value'], d['ms_per_step'], d['e2e'])"; tail -5 gpurun_out/r02_bench11_n2.err
