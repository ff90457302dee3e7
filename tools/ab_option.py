"""Same-process A/B of one engine option on the config-2 shot batch: python tools/ab_option.py prefetch 0 -1
[--shots N] [--reps R].  Prints device-side ms per step (per-pass CUDA events) for each value, interleaved."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dqc_simulator_b200 import experiment  # noqa: E402
from dqc_simulator_b200.engine import Engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("key")
    ap.add_argument("values", nargs="+", type=int)
    ap.add_argument("--shots", type=int, default=100000)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--circuit", default="grover")
    a = ap.parse_args()
    ops, _ = experiment.load_c2(a.circuit)
    dqc = experiment.c2_hardware(noisy=True)
    data, ideal = experiment.ideal_output_ket(ops, lambda f, n, b: Engine(f, n, b))
    eng = Engine("dm", 7, a.shots, seed=2024)
    out = {v: [] for v in a.values}
    fid = {}
    for rep in range(a.reps + 1):
        for v in a.values:
            eng.set_option(a.key, v)
            eng.time_passes(True)
            res = experiment.run_shot_batch(ops, dqc, eng, data, ideal)
            ms, n, _ = eng.pass_time()
            eng.time_passes(False)
            fid[v] = float(res.fidelity.mean())
            if rep:   # first round warms up
                out[v].append(round(ms, 2))
    print(json.dumps({"option": a.key, "shots": a.shots, "pass_ms": out, "mean_fidelity": fid}))


if __name__ == "__main__":
    main()
