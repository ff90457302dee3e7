timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_tile_pass -s 40 -c 3 -f -o gpurun_out/r02r_t11 python bench.py --shots 16384 --steps 1 --warmup 1 --no-cpu-baseline --tile-bits 11 --opt threads=128 > gpurun_out/r02r_ncu11.log 2>&1
echo rc=$?
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_tile_pass -s 40 -c 3 -f -o gpurun_out/r02r_t10 python bench.py --shots 16384 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02r_ncu10.log 2>&1
echo rc=$?
