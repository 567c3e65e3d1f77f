# 8-GPU bench line, launched the way the driver launches it.
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench8.err
echo rc=$?; tail -2 gpurun_out/bench8.err; cut -c1-300 gpurun_out/bench_8gpu.json
