for cfg in "--opt smem_pad=0" "--opt smem_pad=6000" "--opt smem_pad=13000" "--opt smem_pad=24000" "--opt smem_pad=6000 --noise depol+dephase" "--noise depol+dephase"; do
echo "== $cfg"
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline $cfg 2>gpurun_out/r03a_err.log | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l);print(d['value'],d['e2e']['value'],d['ms_per_step'],d['passes_per_step'],d['gpu_launches'],d['roofline']['launches'],d['roofline']['avg_launch_ms'], d['mean_fidelity'])
    else: print(l[:300])"
tail -3 gpurun_out/r03a_err.log
done
