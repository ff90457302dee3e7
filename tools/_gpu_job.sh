set -x
mkdir -p gpurun_out
R=r01t
timeout 600 python -m pytest tests -m gpu -x -q 2>gpurun_out/${R}_pytest.err | tail -8 > gpurun_out/${R}_pytest.log; cat gpurun_out/${R}_pytest.log
for i in 1 2; do
 DQE_LIB=build/libdqcengine_prev.so timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/${R}_ab_prev_$i.json 2>/dev/null
 timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/${R}_ab_new_$i.json 2>/dev/null
done
timeout 300 python tools/bench_c4.py > gpurun_out/${R}_bench_c4.log 2>&1; tail -1 gpurun_out/${R}_bench_c4.log
