set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02o_gputests.log 2>&1
echo tests rc=$?
tail -5 gpurun_out/r02o_gputests.log
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02o_bench.json 2> gpurun_out/r02o_bench.err
echo bench rc=$?
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --opt catfuse=0 > gpurun_out/r02o_bench_nofuse.json 2>> gpurun_out/r02o_bench.err
