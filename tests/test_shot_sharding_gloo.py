"""World-size-2 gloo test of the shot-sharded path: each rank runs its contiguous block
of global shot ids, Philox is keyed by the global id, and one all-reduce merges the results.  The
merged histogram / fidelity must equal a single-process run of all shots.  CPU only (the state
backend here is the oracle; the N > 1 GPU path differs only in the backend object and NCCL)."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from dqc_simulator_b200 import executor, experiment
from oracle.numpy_engine import OracleEngine

TOTAL = 10


def _run(lo, hi):
    ops, _ = experiment.load_c2("ghz")
    dqc = experiment.c2_hardware(noisy=True, F_werner=0.9, meas_error_prob=0.2)
    mk = lambda f, n, b: OracleEngine(f, n, b, seed=0)
    data, ideal = experiment.ideal_output_ket(ops, mk)
    be = OracleEngine("dm", 7, hi - lo, seed=2024, shot_offset=lo)
    ex = executor.DQCExecutor(ops, dqc, be).run()
    axes = ex.axes_of(data)
    fid = np.array([be.fidelity_ket(axes, ideal, s) for s in range(hi - lo)])
    return fid, be.creg.copy()


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = experiment.shard_range(TOTAL, rank, world)
    fid, creg = _run(lo, hi)
    mean_f, n, hist = experiment.all_reduce_results(fid, experiment.outcome_histogram(creg, 4))
    q.put((rank, lo, hi, mean_f, n, hist.tolist(), creg.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_one():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    fid, creg = _run(0, TOTAL)
    assert [g[1:3] for g in got] == [(0, 5), (5, 10)]
    # per-shot outcomes are identical under sharding (bit-exact) ...
    assert got[0][6] + got[1][6] == creg.tolist()
    # ... so the reduced histogram and mean fidelity equal the single-process run
    for g in got:
        assert g[4] == TOTAL
        assert g[5] == experiment.outcome_histogram(creg, 4).tolist()
        assert abs(g[3] - fid.mean()) < 1e-12
    assert len(set(creg.tolist())) > 1  # the trajectories really differ


def test_shard_range_partitions():
    for total in (1, 7, 100000):
        for world in (1, 2, 3, 8):
            r = [experiment.shard_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
