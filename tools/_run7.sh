timeout 900 python -m pytest tests/test_gpu_sharded_multi.py -m gpu -q 2>&1 | tail -3
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02w_bench_x2.json 2> gpurun_out/r02w_bench_x2.err
echo rc=$?
python -c "
import json
d=json.load(open('gpurun_out/r02w_bench_x2.json'))
print(d['value'],d['e2e']['value'],d['n_gpus'],d['ms_per_step'])
s=d.get('sharded_ket',{})
for k in ('weak','strong'):
    if k in s: print(k,{x:s[k][x] for x in ('n','seconds','exchanges','nvlink_GBps_per_dir_per_gpu','pass_GBps','max_abs_err')})
"
tail -3 gpurun_out/r02w_bench_x2.err
