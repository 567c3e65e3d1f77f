# 4-GPU bench line, launched the way the driver launches it.
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_4gpu.json 2> gpurun_out/bench4.err
echo rc=$?; tail -2 gpurun_out/bench4.err; cut -c1-300 gpurun_out/bench_4gpu.json
