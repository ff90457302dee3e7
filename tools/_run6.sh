timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02w_gputests.log 2>&1
echo tests rc=$?
tail -3 gpurun_out/r02w_gputests.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r02w_bench_default.json 2>gpurun_out/r02w_bench.err; python -c "
import json;d=json.load(open('gpurun_out/r02w_bench_default.json'));print(d['value'],d['e2e']['value'],d['roofline']['frac'],d['ms_per_step'],d['mean_fidelity'],d['passes_per_step'],d['cpu_baseline'])"
