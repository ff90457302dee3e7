for cfg in "--tile-bits 11 --opt threads=128 --opt catfuse=0" "--opt catfuse=0" "--tile-bits 11 --opt threads=128 --opt catfuse=1" "--tile-bits 11 --opt threads=128 --circuit qft" "--circuit qft" "--tile-bits 11 --opt threads=128 --noise depol+dephase"; do
  echo "== $cfg"
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline $cfg 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l);print(d['value'],d['e2e']['value'],d['roofline']['frac'],d['ms_per_step'],d['mean_fidelity'],d['passes_per_step'])
    else: print(l[:300])"
done
