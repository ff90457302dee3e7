"""Wall-clock breakdown of one end-to-end shot-batch step (bench.py's e2e leg): where the time between
the device-timed flush and the user-visible result goes.  Usage: python tools/e2e_breakdown.py [--shots N]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dqc_simulator_b200 import executor, experiment  # noqa: E402
from dqc_simulator_b200.engine import Engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shots", type=int, default=100000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--circuit", default="grover")
    a = ap.parse_args()
    ops, _ = experiment.load_c2(a.circuit)
    dqc = experiment.c2_hardware(noisy=True)
    data, ideal = experiment.ideal_output_ket(ops, lambda f, n, b: Engine(f, n, b))
    eng = Engine("dm", 7, a.shots, seed=2024)
    rows = []
    for step in range(a.steps):
        t = [time.perf_counter()]
        eng.reinit(); t.append(time.perf_counter())
        ex = executor.DQCExecutor(ops, dqc, eng).run(); t.append(time.perf_counter())
        axes = ex.axes_of(data)
        eng.flush(); t.append(time.perf_counter())
        eng.sync(); t.append(time.perf_counter())
        fid = eng.fidelity_ket_all(axes, ideal); t.append(time.perf_counter())
        creg = eng.read_creg_words(); t.append(time.perf_counter())
        hist = experiment.outcome_histogram(creg); t.append(time.perf_counter())
        names = ["reinit", "walk", "flush_call", "device_wait", "fidelity", "creg", "hist"]
        row = {n: round(1e3 * (t[i + 1] - t[i]), 2) for i, n in enumerate(names)}
        row["total"] = round(1e3 * (t[-1] - t[0]), 2)
        rows.append(row)
        print(json.dumps(row), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/e2e_breakdown.json", "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
