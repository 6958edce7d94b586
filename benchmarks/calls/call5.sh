set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks5.csv &
SMI=$!
CMD="python benchmarks/ab_kernels.py --rounds 1 --reps 3 --warm 3 --kinds inverse vv_dsp_b200/lib/libvvdsp_b200.so"
$CMD > gpurun_out/plain5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:istft_ws -s 2 -c 1 -o gpurun_out/r02_ws_v2 $CMD > gpurun_out/ncu5.log 2>&1
tail -3 gpurun_out/ncu5.log
python benchmarks/ab_kernels.py --rounds 10 --kinds inverse vv_dsp_b200/lib/libvvdsp_b200_r1.so vv_dsp_b200/lib/libvvdsp_b200.so > gpurun_out/r02_ab5.jsonl
cat gpurun_out/r02_ab5.jsonl
kill $SMI
