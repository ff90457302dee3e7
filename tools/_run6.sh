timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02z_gputests.log 2>&1
echo tests rc=$?
tail -4 gpurun_out/r02z_gputests.log
for cfg in "" "--noise depol+dephase" "--circuit qft" "--shared-timeline"; do
echo "== $cfg"
python bench.py --steps 3 --warmup 2 --no-cpu-baseline $cfg 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l);print(d['value'],d['e2e']['value'],d['roofline']['frac'],d['ms_per_step'],d['mean_fidelity'],d['passes_per_step'])
    else: print(l[:300])"
done
