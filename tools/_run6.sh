set -x
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --circuit qft > gpurun_out/r02w_bench_qft.json 2>>gpurun_out/r02w_bench.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --noise depol+dephase > gpurun_out/r02w_bench_dephase.json 2>>gpurun_out/r02w_bench.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --shared-timeline > gpurun_out/r02w_bench_shared.json 2>>gpurun_out/r02w_bench.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --opt catfuse=0 > gpurun_out/r02w_bench_nofuse.json 2>>gpurun_out/r02w_bench.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --opt catfuse=1 > gpurun_out/r02w_bench_fuse1.json 2>>gpurun_out/r02w_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02w_bench_reference.json 2>>gpurun_out/r02w_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02w_launches.csv python bench.py --shots 16384 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02w_ncu_launches.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_tile_pass -s 170 -c 4 -f -o gpurun_out/r02w_full python bench.py --shots 16384 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02w_ncu.log 2>&1
python tools/microbench.py > gpurun_out/r02w_microbench.json 2>>gpurun_out/r02w_bench.err
ls -la gpurun_out | tail -15
