# 2-GPU lines, launched the way the driver launches them: the headline bench and the mixed corpus (BASELINE configs[4]).
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench2.err
echo rc=$?; tail -2 gpurun_out/bench2.err; cut -c1-400 gpurun_out/bench_2gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 scripts/bench_mixed.py --mib-per-class 1024 --waves 2 > gpurun_out/bench_mixed_2gpu.json 2> gpurun_out/bench_mixed2.err
echo rc=$?; tail -2 gpurun_out/bench_mixed2.err; cat gpurun_out/bench_mixed_2gpu.json
