python bench.py > gpurun_out/r03j_bench_default.json 2>gpurun_out/r03j_bench.err
tail -2 gpurun_out/r03j_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r03j_launches.csv python bench.py --shots 16384 --steps 1 --warmup 1 --no-cpu-baseline --no-other-configs --no-variants > gpurun_out/r03j_ncu_launches.log 2>&1
ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k 'regex:int.2, .int.1, ' --launch-skip 20 --launch-count 3 -f -o gpurun_out/r03j_full python bench.py --shots 16384 --steps 1 --warmup 1 --no-cpu-baseline --no-other-configs --no-variants > gpurun_out/r03j_ncu.log 2>&1
ls -la gpurun_out | tail -5
