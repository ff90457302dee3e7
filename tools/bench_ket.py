"""State-vector configs (BASELINE configs 3 and 5-shape) on one GPU: time, passes, algorithmic GB/s and
gate-updates/s, fused vs one-op-per-pass.  Usage: python tools/bench_ket.py [--reps 3]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dqc_simulator_b200 import executor, hardware, opstream  # noqa: E402
from dqc_simulator_b200.engine import Engine  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CASES = [("c3_qft26_4qpu_cat.json.gz", 28, 4, 9), ("c5_qft30_mono.json.gz", 30, 1, 30),
         ("c5_ghz30_mono.json.gz", 30, 1, 30)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--case", default="", help="substring filter on the fixture name")
    ap.add_argument("--opt", action="append", default=[], help="engine option key=value (repeatable)")
    ap.add_argument("--fused-only", action="store_true")
    ap.add_argument("--ancilla-free", action="store_true",
                    help="ideal hardware: cat-comm primitives as direct remote-controlled gates (executor ancilla_free_cat)")
    a = ap.parse_args()
    out = []
    for name, nq, nqpu, npos in CASES:
        if a.case not in name:
            continue
        ops, _ = opstream.load(os.path.join(GOLDEN, name))
        dqc = hardware.DQCSpec.uniform(nqpu, num_positions=npos, num_comm_qubits=2 if nqpu > 1 else 0)
        for fuse in ((1,) if a.fused_only else (1, 0)):
            af = a.ancilla_free and nqpu > 1
            e = Engine("ket", nq - 2 if af else nq, 1, seed=1)
            e.set_option("fuse", fuse)
            e.set_option("merge", fuse)
            for kv in a.opt:
                k, v = kv.split("=")
                e.set_option(k, int(v))
            best = None
            for rep in range(a.reps + 1):
                e.reinit()
                e.reset_stats()
                e.sync()
                e.time_passes(True)
                t0 = time.perf_counter()
                ex = executor.DQCExecutor(ops, dqc, e, ancilla_free_cat=af).run()
                e.sync()
                wall = time.perf_counter() - t0
                ms, n, amps = e.pass_time()
                e.time_passes(False)
                st = e.stats()
                if rep and (best is None or wall < best["wall_s"]):
                    best = dict(case=name, fused=bool(fuse), opts=a.opt, ancilla_free=bool(af), wall_s=wall, pass_ms=ms, passes=n,
                                api_ops=st["api_ops"], micro_ops=st["micro_ops"],
                                GBps=32.0 * amps / (ms * 1e-3) / 1e9,
                                gate_updates_per_s=st["api_ops"] / wall,
                                amp_updates_per_s=amps / wall)
            print(json.dumps(best), flush=True)
            out.append(best)
            e.close()
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/bench_ket.json", "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
