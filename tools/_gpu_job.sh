set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r01n_pytest.log; cat gpurun_out/r01n_pytest.log
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r01n_bench_new.json 2> gpurun_out/r01n_bench_new.err
timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --circuit qft > gpurun_out/r01n_bench_new_qft.json 2> gpurun_out/r01n_bench_new_qft.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_tile_pass -s 20 -c 3 -f -o gpurun_out/prof_r01n python bench.py --circuit qft --shots 16384 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r01n.log 2>&1
