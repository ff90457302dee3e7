"""State-vector sharding by high-order qubits on CPU: world sizes 2 and 4 over
gloo with a NumPy stand-in for the per-rank engine.  Checks the slot map / exchange logic of
``ShardedKet`` against an unsharded simulation, gate by gate semantics included: diagonal and
control-only uses of a global axis must not communicate, targets on a global axis must."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dqc_simulator_b200 import executor, hardware, opstream
from dqc_simulator_b200.sharded import ShardedKet

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class NumpyShard:
    """Fixed-layout 2^m ket, axis a = bit a (test stand-in for the per-rank CUDA engine)."""

    def __init__(self, m, first_rank):
        self.m = m
        self.t = torch.zeros(2 ** m, dtype=torch.complex128)
        if first_rank:
            self.t[0] = 1
        self.v = self.t.numpy()

    def buffer(self):
        return self.t

    def sync(self):
        pass

    def _apply(self, axes, mat):
        k = len(axes)
        psi = self.v.reshape((2,) * self.m)
        dims = [self.m - 1 - a for a in axes]
        mt = np.asarray(mat, dtype=complex).reshape((2,) * (2 * k))
        out = np.tensordot(mt, psi, axes=(list(range(k, 2 * k)), dims))
        out = np.moveaxis(out, list(range(k)), dims)
        self.v[...] = out.reshape(-1)

    def apply_1q(self, a, m):
        self._apply([a], m)

    def apply_2q(self, a, b, m):
        self._apply([a, b], m)

    def apply_ctrl_1q(self, c, t, m):
        full = np.eye(4, dtype=complex)
        full[2:, 2:] = np.asarray(m).reshape(2, 2)
        self._apply([c, t], full)

    def apply_diag(self, axes, d):
        self._apply(list(axes), np.diag(np.asarray(d, dtype=complex)))

    def apply_kq(self, axes, m):
        self._apply(list(axes), m)

    def probs_diag(self, pairs, target):
        i = np.arange(2 ** self.m)
        keep = np.ones(2 ** self.m, dtype=bool)
        for a, b, fixed in pairs:
            keep &= ((i >> a) & 1) == (fixed if fixed >= 0 else (i >> b) & 1)
        w = np.where(keep, self.v.real, 0.0)
        if target < 0:
            return float(w.sum()), 0.0
        one = ((i >> target) & 1) == 1
        return float(w[~one].sum()), float(w[one].sum())

    def scale(self, s):
        self.v *= s

    def probs(self, axis):
        p = (np.abs(self.v) ** 2).reshape(-1, 2, 2 ** axis).sum(axis=(0, 2))
        return float(p[0]), float(p[1])

    def project(self, axis, outcome, scale, p_local):
        b = self.v.reshape(-1, 2, 2 ** axis)
        b[:, 1 - outcome, :] = 0
        b[:, outcome, :] *= scale
        return 1.0


def rand_u(rng, d):
    q, r = np.linalg.qr(rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d)))
    return q * (np.diag(r) / np.abs(np.diag(r)))


def random_circuit(n, seed, steps=60):
    rng = np.random.default_rng(seed)
    ops = []
    H = np.array([[1, 1], [1, -1]]) / np.sqrt(2)
    for a in range(n):
        ops.append(("1q", a, H))
    cnot = np.eye(4)[[0, 1, 3, 2]]
    for _ in range(steps):
        kind = rng.choice(["1q", "cnot", "2q", "ctrl", "diag1", "diag2", "diag3", "cz"])
        ax = [int(x) for x in rng.choice(n, size=3, replace=False)]
        if kind == "1q":
            ops.append(("1q", ax[0], rand_u(rng, 2)))
        elif kind == "cnot":
            ops.append(("2q", ax[0], ax[1], cnot))
        elif kind == "2q":
            ops.append(("2q", ax[0], ax[1], rand_u(rng, 4)))
        elif kind == "ctrl":
            ops.append(("ctrl", ax[0], ax[1], rand_u(rng, 2)))
        elif kind == "cz":
            ops.append(("2q", ax[0], ax[1], np.diag([1, 1, 1, -1])))
        elif kind == "diag1":
            ops.append(("diag", ax[:1], np.exp(1j * rng.normal(size=2))))
        elif kind == "diag2":
            ops.append(("diag", ax[:2], np.exp(1j * rng.normal(size=4))))
        else:
            ops.append(("diag", ax, np.exp(1j * rng.normal(size=8))))
    return ops


def play(be, ops):
    for op in ops:
        if op[0] == "1q":
            be.apply_1q(op[1], op[2])
        elif op[0] == "2q":
            be.apply_2q(op[1], op[2], op[3])
        elif op[0] == "ctrl":
            be.apply_ctrl_1q(op[1], op[2], op[3])
        else:
            be.apply_diag(op[1], op[2])


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = int(np.log2(world))
    out = {}
    # 1. random circuit
    sk = ShardedKet(n, NumpyShard(n - g, rank == 0), rank, world)
    play(sk, random_circuit(n, 7))
    out["random"] = sk.gather_logical()
    out["exchanges"] = sk.exchanges
    # 2. diagonal / control-only ops on global axes never communicate
    sk2 = ShardedKet(n, NumpyShard(n - g, rank == 0), rank, world)
    rng = np.random.default_rng(3)
    for a in range(n - g):
        sk2.apply_1q(a, rand_u(rng, 2))
    top = n - 1
    sk2.apply_diag([top, 0], np.exp(1j * rng.normal(size=4)))
    sk2.apply_ctrl_1q(top, 1, rand_u(rng, 2))
    sk2.apply_2q(top, 2, np.eye(4)[[0, 1, 3, 2]])        # CNOT with a global control
    out["no_comm_exchanges"] = sk2.exchanges
    sk2.apply_1q(top, rand_u(rng, 2))                     # a target on the global axis must exchange
    out["after_target_exchanges"] = sk2.exchanges
    # 3. the reference-compiled config-5-shaped circuits through the executor
    for name in ("c5_ghz12_mono.json", "c5_qft12_mono.json"):
        ops, _ = opstream.load(os.path.join(GOLDEN, name))
        sk3 = ShardedKet(12, NumpyShard(12 - g, rank == 0), rank, world)
        executor.DQCExecutor(ops, hardware.DQCSpec.uniform(1, num_positions=12, num_comm_qubits=0), sk3).run()
        out[name] = sk3.gather_logical()
        out[name + "_norm"] = sk3.norm2()
    # 4. measurement across shards: same outcome everywhere, state collapsed and normalised
    sk4 = ShardedKet(n, NumpyShard(n - g, rank == 0), rank, world, seed=5)
    play(sk4, random_circuit(n, 11, steps=20))
    out["m_global"] = sk4.measure(n - 1)
    out["m_local"] = sk4.measure(0)
    out["measured"] = sk4.gather_logical()
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def noisy_program(be, n, seed):
    """Gates, channels and measurements of a noisy density-matrix run, through the common backend surface."""
    rng = np.random.default_rng(seed)
    H = np.array([[1, 1], [1, -1]]) / np.sqrt(2)
    for a in range(n):
        be.alloc_axis(a)
        be.apply_1q(a, H)
    outcomes = []
    for step in range(40):
        kind = rng.choice(["1q", "cnot", "2q", "ctrl", "diag2", "dep1", "dep2", "pauli", "damp"])
        ax = [int(x) for x in rng.choice(n, size=2, replace=False)]
        if kind == "1q":
            be.apply_1q(ax[0], rand_u(rng, 2))
        elif kind == "cnot":
            be.apply_2q(ax[0], ax[1], np.eye(4)[[0, 1, 3, 2]])
        elif kind == "2q":
            be.apply_2q(ax[0], ax[1], rand_u(rng, 4))
        elif kind == "ctrl":
            be.apply_ctrl_1q(ax[0], ax[1], rand_u(rng, 2))
        elif kind == "diag2":
            be.apply_diag(ax, np.exp(1j * rng.normal(size=4)))
        elif kind == "dep1":
            be.depolarize([ax[0]], float(rng.uniform(0.01, 0.3)))
        elif kind == "dep2":
            be.depolarize(ax, float(rng.uniform(0.01, 0.3)))
        elif kind == "pauli":
            p = rng.dirichlet([8, 1, 1, 1])
            be.pauli_channel(ax[0], [float(x) for x in p])
        else:
            g = float(rng.uniform(0.05, 0.4))
            be.kraus_1q(ax[0], [np.array([[1, 0], [0, np.sqrt(1 - g)]]), np.array([[0, np.sqrt(g)], [0, 0]])])
        if step in (15, 30):
            m = int(step == 15)
            outcomes.append(be.measure(ax[0], 2, m, False, 0.0))
            be.set_predicate(2)
            be.apply_1q(ax[1], np.array([[0, 1], [1, 0]]))           # runs only after the outcome-1 measurement
            be.set_predicate(-1)
    return outcomes


def _dm_worker(rank, world, port, n, q):
    from dqc_simulator_b200.sharded import ShardedDM

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = int(np.log2(world))
    sd = ShardedDM(n, NumpyShard(2 * n - g, rank == 0), rank, world, seed=5)
    noisy_program(sd, n, 21)
    q.put((rank, dict(rho=sd.gather_rho(), trace=sd.trace(), exchanges=sd.exchanges)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_density_matrix_sharded_by_row_bits_matches_oracle(world):
    """SURVEY 8e row 3 / VERDICT e3: ONE rho sharded over the ranks by its high-order row bits (ShardedDM): gates,
    depolarising (1 and 2 qubits), Pauli and Kraus channels, forced measurements with a predicated correction --
    the gathered matrix equals the oracle's density-matrix run of the same program (n = 6: 4096 elements)."""
    from oracle.numpy_engine import OracleEngine

    n = 6
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dm_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    o = OracleEngine("dm", n, 1, seed=5)
    noisy_program(o, n, 21)
    want = o.full_state(0)[0]
    for r in range(world):
        assert np.abs(got[r]["rho"] - want).max() < 1e-12
        assert abs(got[r]["trace"] - 1.0) < 1e-12
        assert got[r]["exchanges"] > 0


def _reference(n, ops):
    be = NumpyShard(n, True)
    play(be, ops)
    return be.v.copy()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_state_vector_matches_unsharded(world):
    from oracle.numpy_engine import OracleEngine

    n = 8
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = _reference(n, random_circuit(n, 7))
    for r in range(world):
        assert np.abs(got[r]["random"] - want).max() < 1e-12
        assert got[r]["exchanges"] > 0
        assert got[r]["no_comm_exchanges"] == 0
        assert got[r]["after_target_exchanges"] == 1
    for name in ("c5_ghz12_mono.json", "c5_qft12_mono.json"):
        ops, _ = opstream.load(os.path.join(GOLDEN, name))
        o = OracleEngine("ket", 12, 1)
        executor.DQCExecutor(ops, hardware.DQCSpec.uniform(1, num_positions=12, num_comm_qubits=0), o).run()
        for r in range(world):
            assert np.abs(got[r][name] - o.full_state(0)[0]).max() < 1e-12
            assert abs(got[r][name + "_norm"] - 1) < 1e-12
    # measurement: identical outcomes on all ranks; result = projected + renormalised reference
    ref = _reference(n, random_circuit(n, 11, steps=20))
    mg, ml = got[0]["m_global"], got[0]["m_local"]
    assert all(got[r]["m_global"] == mg and got[r]["m_local"] == ml for r in range(world))
    idx = np.arange(2 ** n)
    keep = (((idx >> (n - 1)) & 1) == mg) & ((idx & 1) == ml)
    proj = np.where(keep, ref, 0)
    proj = proj / np.linalg.norm(proj)
    assert np.abs(got[0]["measured"] - proj).max() < 1e-12
