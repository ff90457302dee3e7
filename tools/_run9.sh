timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
DQE_PRINT_PASS_MS=1 python tools/bench_ket.py --case qft30 --fused-only --reps 1 --check 2>&1 | tail -11
python tools/bench_ket.py --case qft30 --fused-only --reps 2 --opt phase_tables=0 2>&1 | tail -1
