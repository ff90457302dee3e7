timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02s_gputests.log 2>&1
echo tests rc=$?
tail -3 gpurun_out/r02s_gputests.log
for cfg in "" "--tile-bits 11 --opt threads=128" "--tile-bits 11 --opt threads=128 --opt pass_ops=30" "--tile-bits 11 --opt threads=128 --opt pass_ops=18" "--opt pass_ops=18" "--tile-bits 11 --opt threads=64 --opt pass_ops=18" "--circuit qft" "--circuit qft --opt pass_ops=18" "--tile-bits 11 --opt threads=128 --circuit qft --opt pass_ops=18"; do
  echo "== $cfg"
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline $cfg 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l);print(d['value'],d['e2e']['value'],d['roofline']['frac'],d['ms_per_step'],d['mean_fidelity'],d['passes_per_step'])
    else: print(l[:300])"
done
