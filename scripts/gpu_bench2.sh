# 2-GPU bench line, launched the way the driver launches it.
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench2.err
echo rc=$?; tail -2 gpurun_out/bench2.err; cut -c1-600 gpurun_out/bench_2gpu.json
