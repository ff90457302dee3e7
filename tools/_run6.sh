timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02u_gputests.log 2>&1
echo tests rc=$?
tail -3 gpurun_out/r02u_gputests.log
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02u_bench.json 2>gpurun_out/r02u_bench.err; python -c "
import json;d=json.load(open('gpurun_out/r02u_bench.json'));print(d['value'],d['e2e']['value'],d['roofline']['frac'],d['ms_per_step'],d['mean_fidelity'],d['passes_per_step'])"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_tile_pass -s 170 -c 3 -f -o gpurun_out/r02u_full python bench.py --shots 16384 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02u_ncu.log 2>&1
echo rc=$?
