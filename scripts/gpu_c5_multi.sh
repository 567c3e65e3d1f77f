# BASELINE configs[4] on N GPUs of one box: bash scripts/gpu_c5_multi.sh N TOTAL_GIB
set -x
N=$1; TOT=$2
mkdir -p gpurun_out
timeout 2400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --config 5 --steps 3 --warmup 3 --total-gib $TOT --wave-gib 2 --no-cpu > gpurun_out/bench_c5_${N}gpu.json 2> gpurun_out/bench_c5_${N}gpu.err; echo rc=$?
tail -3 gpurun_out/bench_c5_${N}gpu.err
cut -c1-600 gpurun_out/bench_c5_${N}gpu.json
