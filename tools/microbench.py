"""Per-op pass timings on the GPU (CUDA events inside the library): achieved algorithmic GB/s
of k_tile_pass for single-op passes, ket and dm.  Usage: python tools/microbench.py [--tile T]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dqc_simulator_b200.engine import Engine  # noqa: E402

H = np.array([[1, 1], [1, -1]], dtype=complex) / np.sqrt(2)
X = np.array([[0, 1], [1, 0]], dtype=complex)
CNOT = np.eye(4, dtype=complex)[[0, 1, 3, 2]]


def rand_u(rng, d):
    q, r = np.linalg.qr(rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d)))
    return q * (np.diag(r) / np.abs(np.diag(r)))


def timed(e, fn, reps=5):
    fn(); e.sync()
    e.time_passes(True)
    for _ in range(reps):
        fn()
    e.sync()
    ms, n, amps = e.pass_time()
    e.time_passes(False)
    return ms, n, amps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tile", type=int, default=None)
    ap.add_argument("--ket-n", type=int, default=28)
    ap.add_argument("--dm-n", type=int, default=7)
    ap.add_argument("--dm-batch", type=int, default=16384)
    a = ap.parse_args()
    rng = np.random.default_rng(0)
    out = []

    def report(name, ms, n, amps, bytes_per_amp=32.0):
        # algorithmic bytes: 32 B per live amplitude per pass (16 read + 16 written); a measurement is two passes,
        # the first of which only READS (dm: the diagonal only) -- credited 16 B / amplitude at most
        gbs = bytes_per_amp * amps / (ms * 1e-3) / 1e9
        out.append(dict(case=name, ms_per_pass=ms / n, passes=n, GBps=gbs))
        print(f"{name:40s} {ms / n:9.4f} ms/pass  {gbs:8.1f} GB/s  ({n} passes)", flush=True)

    # ---- ket, single shot
    n = a.ket_n
    e = Engine("ket", n, 1, tile_bits=a.tile, fuse=False)
    e.set_option("merge", 0)          # every API op = one pass (no host-side folding of repeated gates)
    for q in range(n):
        e.alloc_axis(q)
        e.apply_1q(q, H)
    e.sync()
    u = rand_u(rng, 2)
    u4 = rand_u(rng, 4)
    for q in (0, 3, 7, 11, 12, 20, n - 1):
        report(f"ket{n} 1q dense q={q}", *timed(e, lambda: e.apply_1q(q, u)))
    for c, t in ((0, 1), (n - 1, 0), (5, n - 1), (n - 2, n - 1)):
        report(f"ket{n} cnot c={c} t={t}", *timed(e, lambda: e.apply_2q(c, t, CNOT)))
    for q0, q1 in ((0, 1), (3, 20), (n - 2, n - 1)):
        report(f"ket{n} 2q dense ({q0},{q1})", *timed(e, lambda: e.apply_2q(q0, q1, u4)))
    report(f"ket{n} diag2 (2,{n-1})", *timed(e, lambda: e.apply_diag([2, n - 1], np.exp(1j * np.arange(4)))))
    e.set_option("fuse", 1)
    e.set_option("merge", 1)

    def many():
        for q in range(8):
            e.apply_1q(q, u)
    report(f"ket{n} 8x 1q fused (q0-7)", *timed(e, many))

    def many2():
        for q in range(n - 4, n):
            e.apply_1q(q, u)
    report(f"ket{n} 4x 1q fused (top 4)", *timed(e, many2))

    # dense k-qubit block (OP_KBLOCK, FP64 tensor cores): 24 gates inside bits 0-3, planner-fused vs gate by gate
    us = [rand_u(rng, 2) for _ in range(16)] + [rand_u(rng, 4) for _ in range(8)]

    def dense4():
        for i, m in enumerate(us):
            if m.shape[0] == 2:
                e.apply_1q(i % 4, m)
            else:
                e.apply_2q(i % 4, (i + 1) % 4, m)
    for kb in (1, 0):
        e.set_option("kblock", kb)
        report(f"ket{n} 24 gates on bits 0-3, kblock={kb}", *timed(e, dense4))
    e.set_option("kblock", 1)
    u16 = rand_u(rng, 16)
    report(f"ket{n} dense 4q block (0,1,2,3)", *timed(e, lambda: e.apply_kq([0, 1, 2, 3], u16)))
    report(f"ket{n} dense 4q block (3,9,17,{n-1})", *timed(e, lambda: e.apply_kq([3, 9, 17, n - 1], u16)))
    u8 = rand_u(rng, 8)
    report(f"ket{n} dense 3q block (0,1,2)", *timed(e, lambda: e.apply_kq([0, 1, 2], u8)))
    e.close()

    # ---- dm, batched
    n, B = a.dm_n, a.dm_batch
    e = Engine("dm", n, B, tile_bits=a.tile, fuse=False)
    e.set_option("merge", 0)
    for q in range(n):
        e.alloc_axis(q)
        e.apply_1q(q, H)
    e.sync()
    for q in (0, 3, n - 1):
        report(f"dm{n}x{B} 1q dense q={q}", *timed(e, lambda: e.apply_1q(q, u)))
        report(f"dm{n}x{B} depol1 q={q}", *timed(e, lambda: e.depolarize([q], 0.01)))
    for c, t in ((0, 1), (n - 1, 0), (n - 2, n - 1)):
        report(f"dm{n}x{B} cnot c={c} t={t}", *timed(e, lambda: e.apply_2q(c, t, CNOT)))
        report(f"dm{n}x{B} depol2 ({c},{t})", *timed(e, lambda: e.depolarize([c, t], 0.01)))
    report(f"dm{n}x{B} 2q dense (1,{n-1})", *timed(e, lambda: e.apply_2q(1, n - 1, u4)))
    report(f"dm{n}x{B} measure q={n-1} keep (probs pass reads only + collapse pass: <= 24 B / amp / pass on average)",
           *timed(e, lambda: e.measure(n - 1, 0, -1, False, 0.003)), bytes_per_amp=24.0)
    e.set_option("fuse", 1)
    e.set_option("merge", 1)

    def seq():
        e.depolarize([0], 1e-3); e.depolarize([1], 1e-3)
        e.apply_2q(0, 1, CNOT); e.depolarize([0, 1], 1e-3)
        e.apply_1q(0, u); e.depolarize([0], 1e-3)
    report(f"dm{n}x{B} 6-op noisy cnot+1q fused", *timed(e, seq))
    e.close()
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/microbench.json", "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
