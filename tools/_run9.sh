ncu --set full --import-source on --clock-control none -k regex:k_tile_pass --launch-skip 3 --launch-count 2 -f -o gpurun_out/r03h_qft30 python tools/bench_ket.py --case qft30 --fused-only --reps 1 > gpurun_out/r03h_ncu.log 2>&1
tail -4 gpurun_out/r03h_ncu.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
