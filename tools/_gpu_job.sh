set -x
mkdir -p gpurun_out
R=r01u
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544"
timeout 240 $TR bench.py --gpus 8 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/${R}_bench_x8.json 2> gpurun_out/${R}_bench_x8.err; tail -c 300 gpurun_out/${R}_bench_x8.json
