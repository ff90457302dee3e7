"""CUDA engine vs oracle, through the C-ABI, on identical seeded op streams.

Bar (BASELINE.json north_star): amplitudes / rho within 1e-10 max-abs (complex128); classical
register contents, live-axis bookkeeping and forced-outcome measurements bit-exact.
"""
import numpy as np
import pytest

from oracle.numpy_engine import OracleEngine

pytestmark = pytest.mark.gpu

TOL = 1e-10


def _engine(*a, **k):
    from dqc_simulator_b200.engine import Engine

    return Engine(*a, **k)


def rand_unitary(rng, d):
    z = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d))
    q, r = np.linalg.qr(z)
    return q * (np.diag(r) / np.abs(np.diag(r)))


X = np.array([[0, 1], [1, 0]], dtype=complex)
Z = np.diag([1, -1]).astype(complex)
H = np.array([[1, 1], [1, -1]], dtype=complex) / np.sqrt(2)
CNOT = np.eye(4, dtype=complex)[[0, 1, 3, 2]]
SWAP = np.eye(4, dtype=complex)[[0, 2, 1, 3]]


def werner(F):
    s = 2 ** -0.5
    b = [np.array(v, dtype=complex) for v in ([s, 0, 0, s], [0, s, s, 0], [s, 0, 0, -s], [0, s, -s, 0])]
    out = F * np.outer(b[0], b[0].conj())
    for v in b[1:]:
        out = out + (1 - F) / 3 * np.outer(v, v.conj())
    return out


class Both:
    """Drives the CUDA engine and the oracle in lock step."""

    def __init__(self, formalism, n, batch, seed, **kw):
        self.g = _engine(formalism, n, batch, seed=seed, shot_offset=11, **kw)
        self.o = OracleEngine(formalism, n, batch, seed=seed, shot_offset=11)
        self.n = n

    def __getattr__(self, name):
        gf, of = getattr(self.g, name), getattr(self.o, name)

        def call(*a, **k):
            of(*a, **k)
            return gf(*a, **k)

        return call

    def check(self, what=""):
        gs, os_ = self.g.full_state(), self.o.full_state()
        err = float(np.abs(gs - os_).max())
        assert err < TOL, f"{what}: max-abs {err:.3e}"
        assert sorted(self.g.live) == sorted(self.o.live)
        assert np.array_equal(self.g.read_creg_words(), self.o.creg), what
        for r in sorted(getattr(self, "regs_written", ())):
            a, b = self.g.read_time_reg(r), self.o.read_time_reg(r)
            assert np.allclose(a, b, rtol=1e-12, atol=1e-15), f"{what}: time register {r}"   # exp(): last ulps
        return err


class ClassicalState:
    """Bookkeeping that keeps a random program inside the contract of dqe_classical_op: a register / creg
    bit is not overwritten while an op queued earlier in the same span (since the last measurement or
    fence) still reads it."""

    def __init__(self, b):
        self.b = b
        b.regs_written = set()
        self.read_regs, self.read_bits = set(), set()
        self.times, self.qs = [], []          # registers holding times / probabilities
        self.scratch_bits = []

    def new_span(self):
        self.read_regs.clear()
        self.read_bits.clear()

    def fresh_reg(self):
        for r in range(128):
            if r not in self.read_regs and r not in self.times[-6:] and r not in self.qs[-6:]:
                self.b.regs_written.add(r)
                return r
        self.b.classical_fence()
        self.new_span()
        return self.fresh_reg()

    def fresh_bit(self):
        for bit in range(40, 64):
            if bit not in self.read_bits:
                return bit
        self.b.classical_fence()
        self.new_span()
        return self.fresh_bit()


def random_program(b, rng, steps, formalism, check_every=0, classical=False):
    n = b.n
    cbit = 0
    cs = ClassicalState(b) if classical else None
    for step in range(steps):
        live = list(b.o.live)
        free = [a for a in range(n) if a not in live]
        choices = []
        if free:
            choices += ["alloc"] * 2
        if len(free) >= 2:
            choices += ["inject"] * 2
        if live:
            choices += ["1q"] * 4 + ["diag1", "pauli", "depol1", "measure", "measure", "x", "h", "pred1q"]
            if formalism == "dm":
                choices += ["kraus"]
        if len(live) >= 2:
            choices += ["2q"] * 2 + ["cnot"] * 3 + ["ctrl", "cz", "diag2", "depol2", "depol2", "swap"]
        if len(live) >= 3:
            choices += ["free", "reset"]
            if formalism == "ket":
                choices += ["diag3"]
        if classical:
            choices += ["cl_time"] * 3 + ["cl_bit"]
            if live and cs.qs:
                choices += ["dyn"] * 4
        c = rng.choice(choices)
        pick = lambda k: [int(x) for x in rng.choice(live, size=k, replace=False)]
        if c == "alloc":
            b.alloc_axis(int(rng.choice(free)))
        elif c == "inject":
            a0, a1 = [int(x) for x in rng.choice(free, size=2, replace=False)]
            kind = rng.integers(3)
            if kind == 0:
                b.inject_2q(a0, a1, np.array([1, 0, 0, 1]) / np.sqrt(2), False)
            elif kind == 1:
                v = rng.normal(size=4) + 1j * rng.normal(size=4)
                b.inject_2q(a0, a1, v / np.linalg.norm(v), False)
            else:
                b.inject_2q(a0, a1, werner(0.9), True)
        elif c == "1q":
            b.apply_1q(pick(1)[0], rand_unitary(rng, 2))
        elif c == "x":
            b.apply_1q(pick(1)[0], X)
        elif c == "h":
            b.apply_1q(pick(1)[0], H)
        elif c == "diag1":
            b.apply_1q(pick(1)[0], np.diag(np.exp(1j * rng.normal(size=2))))
        elif c == "cl_time":
            # a few instructions of clock arithmetic ending in a channel strength q = 1 - exp(-dt * rate)
            from dqc_simulator_b200 import timeline as tl

            t0 = cs.fresh_reg()
            b.classical_op(tl.CL_CONST, t0, imm=float(rng.uniform(0, 1e6)))
            cs.times.append(t0)
            t1 = cs.fresh_reg()
            if cbit and rng.integers(2):
                bit = int(rng.integers(0, min(cbit, 40)))
                tt = cs.fresh_reg()
                b.classical_op(tl.CL_ADDI, tt, t0, imm=135e3)
                cs.times.append(tt)
                b.classical_op(tl.CL_SEL, t1, tt, t0, cbit=bit)
                cs.read_bits.add(bit)
            else:
                src = int(rng.choice(cs.times))
                b.classical_op(tl.CL_MAX, t1, t0, src)
                cs.read_regs.add(src)
            cs.times.append(t1)
            kind = int(rng.integers(4))
            t2 = cs.fresh_reg()
            if kind == 0:
                b.classical_op(tl.CL_ADDI, t2, t1, imm=6e5)
            elif kind == 1:
                b.classical_op(tl.CL_CEILQ, t2, t1, imm=6e5)
            elif kind == 2:
                b.classical_op(tl.CL_FMAI, t2, t1, t0, imm=3.0)
            else:
                b.classical_op(tl.CL_RSUBI, t2, t1, imm=5e6)
            cs.times.append(t2)
            dt = cs.fresh_reg()
            b.classical_op(tl.CL_SUB, dt, t2, t0)
            cs.times.append(dt)
            q = cs.fresh_reg()     # QEXP results feed channels only (dqe_classical_op contract)
            b.classical_op(tl.CL_QEXP, q, dt, imm=float(rng.uniform(1e-8, 3e-7)))
            cs.qs.append(q)
            cs.read_regs.update((t0, t1, t2, dt))
        elif c == "cl_bit":
            from dqc_simulator_b200 import timeline as tl

            srcs = [int(x) for x in rng.integers(0, max(min(cbit, 40), 1), size=2)] + cs.scratch_bits[-2:]
            a0, a1 = int(rng.choice(srcs)), int(rng.choice(srcs))
            d = cs.fresh_bit()
            b.classical_op(int(rng.choice([tl.CL_BAND, tl.CL_BANDN, tl.CL_BOR, tl.CL_BNOT, tl.CL_BCOPY])), d, a0, a1)
            cs.read_bits.update((a0, a1))
            cs.scratch_bits.append(d)
        elif c == "dyn":
            q = int(rng.choice(cs.qs[-4:]))
            kinds = [0, 1, 2, 3] if formalism == "dm" else [0, 1, 3]
            if cs.scratch_bits and rng.integers(3) == 0:
                bit = cs.scratch_bits[-1]
                b.set_predicate(bit)
                b.dyn_channel(pick(1)[0], int(rng.choice(kinds)), q)
                b.set_predicate(-1)
                cs.read_bits.add(bit)
            else:
                b.dyn_channel(pick(1)[0], int(rng.choice(kinds)), q)
            cs.read_regs.add(q)
        elif c == "pred1q":
            bit = int(rng.integers(0, max(cbit, 1))) % 64
            if cs:
                bit = bit % 40
                cs.read_bits.add(bit)
            if rng.integers(2):
                b.cond_1q(pick(1)[0], X if rng.integers(2) else Z, bit)
            else:
                b.set_predicate(bit)
                b.apply_1q(pick(1)[0], rand_unitary(rng, 2))
                if len(live) >= 2:
                    b.apply_2q(*pick(2), CNOT)
                    b.depolarize(pick(2), 0.05)
                b.set_predicate(-1)
        elif c == "pauli":
            p = rng.dirichlet([8, 1, 1, 1])
            b.pauli_channel(pick(1)[0], p)
        elif c == "depol1":
            b.depolarize(pick(1), float(rng.uniform(0, 0.3)))
        elif c == "depol2":
            b.depolarize(pick(2), float(rng.uniform(0, 0.3)))
        elif c == "kraus":
            g = float(rng.uniform(0, 0.5))
            b.kraus_1q(pick(1)[0], [[[1, 0], [0, np.sqrt(1 - g)]], [[0, np.sqrt(g)], [0, 0]]])
        elif c == "2q":
            b.apply_2q(*pick(2), rand_unitary(rng, 4))
        elif c == "cnot":
            b.apply_2q(*pick(2), CNOT)
        elif c == "swap":
            b.apply_2q(*pick(2), SWAP)
        elif c == "cz":
            b.apply_2q(*pick(2), np.diag([1, 1, 1, -1]).astype(complex))
        elif c == "ctrl":
            b.apply_ctrl_1q(*pick(2), rand_unitary(rng, 2))
        elif c == "diag2":
            b.apply_diag(pick(2), np.exp(1j * rng.normal(size=4)))
        elif c == "diag3":
            b.apply_diag(pick(3), np.exp(1j * rng.normal(size=8)))
        elif c == "measure":
            forced = -1 if rng.integers(4) else int(rng.integers(2))
            b.measure(pick(1)[0], cbit % (40 if cs else 64), forced, bool(rng.integers(2)),
                      float(rng.choice([0.0, 0.0, 0.05])))
            cbit += 1
            if cs:
                cs.new_span()
        elif c == "free":
            b.free_axis(pick(1)[0])
        elif c == "reset":
            b.reset(pick(1)[0])
        if check_every and (step + 1) % check_every == 0:
            b.check(f"step {step} ({c})")
    b.check("end")


@pytest.mark.parametrize("formalism,n,batch", [("ket", 6, 5), ("dm", 4, 5), ("ket", 10, 3), ("dm", 6, 2)])
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_programs_stepwise(formalism, n, batch, seed):
    rng = np.random.default_rng(1000 * seed + n)
    b = Both(formalism, n, batch, seed)
    random_program(b, rng, 60, formalism, check_every=1)


@pytest.mark.parametrize("formalism,n,batch,tile", [("ket", 12, 2, 8), ("ket", 14, 1, 10), ("dm", 7, 2, 9),
                                                     ("dm", 7, 3, 12), ("ket", 9, 4, 5), ("dm", 5, 3, 6)])
@pytest.mark.parametrize("fuse", [True, False])
def test_random_programs_fused(formalism, n, batch, tile, fuse):
    """Long queues, one flush: exercises the pass planner (tiles smaller than the register, ops
    with target bits outside the contiguous run, controls / diagonals on outer bits)."""
    rng = np.random.default_rng(77 + n + tile)
    b = Both(formalism, n, batch, 5, tile_bits=tile, fuse=fuse)
    random_program(b, rng, 150, formalism, check_every=50)
    st = b.g.stats()
    assert st["passes"] > 0 and st["micro_ops"] >= st["passes"]
    if fuse:
        assert st["micro_ops"] > st["passes"]


def cat_like_program(b, rng, rounds, classical=False):
    """Remote-gate-shaped programs (dqc_control.py:316-403): a fresh pair (x, y), a short run of constant ops on
    (A, x) -- CNOT either way, controlled-U, gate noise, one-qubit gates and channels on x or A --, then x measured
    and discarded; unrelated ops on other axes and predicated corrections in between.  What
    fuse_cat_measure (dqe_engine.cu) rewrites into one fused measurement, and what it must leave alone."""
    from dqc_simulator_b200 import timeline as tl

    n = b.n
    data = list(range(n - 2))
    for a in data:
        b.alloc_axis(a)
        b.apply_1q(a, rand_unitary(rng, 2))
    for i in range(len(data) - 1):
        b.apply_2q(data[i], data[i + 1], CNOT)
    cbit = 0
    for r in range(rounds):
        x, y = (n - 1, n - 2) if rng.integers(2) else (n - 2, n - 1)
        A, other = [int(v) for v in rng.choice(data, size=2, replace=False)]
        kind = rng.integers(3)
        first, second = (x, y) if rng.integers(2) else (y, x)
        if kind == 0:
            b.inject_2q(first, second, werner(float(rng.uniform(0.7, 1.0))), True)
        elif kind == 1:
            v = rng.normal(size=4) + 1j * rng.normal(size=4)
            b.inject_2q(first, second, v / np.linalg.norm(v), False)
        else:
            b.inject_2q(first, second, np.array([1, 0, 0, 1]) / np.sqrt(2), False)
        if classical:
            b.classical_op(tl.CL_CONST, 2 * (r % 8), imm=float(rng.uniform(1e5, 1e7)))
            b.classical_qexp2(2 * (r % 8) + 1, 2 * (r % 8), 255, 0.0, float(rng.uniform(1e-9, 1e-7)))
            b.dyn_channel(A, 0, 2 * (r % 8) + 1)          # per-shot idle noise in front of the interaction: stays outside
            b.regs_written.update((2 * (r % 8), 2 * (r % 8) + 1))
        if rng.integers(3) == 0:
            b.apply_1q(other, rand_unitary(rng, 2))       # unrelated, commutes
        spoil = rng.integers(8) == 0                        # now and then something the rewrite must refuse
        for _ in range(int(rng.integers(1, 5))):
            c = rng.choice(["cx", "xc", "ctrl", "depol2", "1qx", "paulix", "1qA", "depolA", "hx"])
            if c == "cx":
                b.apply_2q(A, x, CNOT)
            elif c == "xc":
                b.apply_2q(x, A, CNOT)
            elif c == "ctrl":
                b.apply_ctrl_1q(*((A, x) if rng.integers(2) else (x, A)), rand_unitary(rng, 2))
            elif c == "depol2":
                b.depolarize([A, x] if rng.integers(2) else [x, A], float(rng.uniform(0, 0.2)))
            elif c == "1qx":
                b.apply_1q(x, rand_unitary(rng, 2))
            elif c == "hx":
                b.apply_1q(x, H)
            elif c == "paulix":
                b.pauli_channel(x, rng.dirichlet([8, 1, 1, 1]))
            elif c == "1qA":
                b.apply_1q(A, rand_unitary(rng, 2))
            else:
                b.depolarize([A], float(rng.uniform(0, 0.2)))
        if spoil:
            c = rng.integers(3)
            if c == 0:
                b.apply_2q(A, other, CNOT)                 # A meets a third axis behind the first (A, x) op
            elif c == 1:
                b.apply_1q(y, H)                           # the partner is touched before the measurement
            else:
                b.apply_2q(y, x, CNOT)
        forced = -1 if rng.integers(4) else int(rng.integers(2))
        b.measure(x, cbit % 40, forced, True, float(rng.choice([0.0, 0.003, 0.05])))
        b.cond_1q(y, X, cbit % 40)
        cbit += 1
        if classical and rng.integers(3):
            # per-shot idle noise on the surviving half between the two measurements: folds into the whole-gate op
            # by linearity in q (depolarising / dephasing / Z flip); amplitude damping is not linear and must not
            kind = int(rng.choice([tl.DYN_DEPOL, tl.DYN_DEPOL, tl.DYN_DEPHASE, tl.DYN_ZFLIP, tl.DYN_AMPDAMP]))
            b.dyn_channel(y, kind, 2 * (r % 8) + 1)
            if rng.integers(4) == 0:
                b.pauli_channel(y, rng.dirichlet([8, 1, 1, 1]))
        # second half: y interacts with another data qubit and is measured out as well
        b.apply_2q(y, other, CNOT)
        b.depolarize([y, other], 0.01)
        b.apply_1q(y, H)
        b.measure(y, cbit % 40, -1, True, 0.003)
        b.cond_1q(A, Z, cbit % 40)
        cbit += 1
        if (r + 1) % 4 == 0:
            b.check(f"round {r}")
    b.check("end")


@pytest.mark.parametrize("n,batch,tile", [(7, 5, 10), (6, 4, 8), (7, 3, 12), (5, 6, 10)])
@pytest.mark.parametrize("classical", [False, True])
def test_fused_inject_measure_runs_keep_parity(n, batch, tile, classical):
    """inject -> interact -> measure-and-discard runs are rewritten at flush into one fused measurement whose axis x
    never goes live, measure-out runs likewise, and whole remote gates into two ops on the data pair
    (fuse_cat_measure, catfuse = 3): same state, classical bits and live axes as the oracle running the ops one by one,
    and as the engine with the rewrite switched off."""
    rng = np.random.default_rng(31 + n + tile)
    b = Both("dm", n, batch, 17, tile_bits=tile)
    if classical:
        b.regs_written = set()
    cat_like_program(b, rng, 12, classical)
    st = b.g.stats()
    rng = np.random.default_rng(31 + n + tile)
    b2 = Both("dm", n, batch, 17, tile_bits=tile)
    b2.g.set_option("catfuse", 0)
    if classical:
        b2.regs_written = set()
    cat_like_program(b2, rng, 12, classical)
    assert np.abs(b.g.full_state() - b2.g.full_state()).max() < TOL
    assert np.array_equal(b.g.read_creg_words(), b2.g.read_creg_words())
    if n - tile // 2 <= 2:   # cluster-fused measurements: the rewrite is active
        assert st["merged_ops"] > b2.g.stats()["merged_ops"]


@pytest.mark.parametrize("opts", [dict(cxperm=0), dict(factor=0), dict(tmap=0), dict(prefetch=0), dict(prefetch=3),
                                  dict(cxperm=0, factor=0, tmap=0), dict(cluster=0), dict(group=0), dict(tma=0), dict(catfuse=0), dict(catfuse=1), dict(catfuse=2),
                                  dict(chain=1), dict(chain=1, tma=0), dict(chain=1, chain_waves=0, chain_max=3)])
@pytest.mark.parametrize("n,batch,tile", [(7, 3, 10), (6, 4, 8), (7, 2, 12)])
def test_dm_kernel_options_keep_parity(n, batch, tile, opts):
    """Every planner / kernel switch of the density-matrix path (X / CX by address permutation, factored
    superoperators inside group sweeps, tensor-map tile copies, L2 prefetch, cluster-fused measurement,
    group sweeps, TMA) gives the oracle's state and classical bits, on or off."""
    rng = np.random.default_rng(4242 + n + tile)
    b = Both("dm", n, batch, 9, tile_bits=tile)
    for k, v in opts.items():
        b.g.set_option(k, v)
    random_program(b, rng, 120, "dm", check_every=40)


@pytest.mark.parametrize("formalism,n,batch,tile", [("dm", 7, 3, 10), ("dm", 6, 4, 8), ("dm", 7, 2, 12), ("dm", 4, 5, 11),
                                                     ("ket", 12, 3, 8), ("ket", 9, 4, 11)])
@pytest.mark.parametrize("opts", [dict(), dict(cluster=0), dict(fuse=0), dict(chain=1), dict(pass_ops=8, chain=1)])
def test_classical_machine_and_dynamic_channels(formalism, n, batch, tile, opts):
    """Per-shot classical registers (dqe_classical_op): clock arithmetic, selects on creg bits, derived bits,
    channels whose strength is read from a register -- with the instructions hoisted behind the previous
    measurement by the planner, in cluster-fused passes, in finish-kernel passes and one op per pass."""
    rng = np.random.default_rng(99 + n + tile)
    b = Both(formalism, n, batch, 13, tile_bits=tile)
    for k, v in opts.items():
        b.g.set_option(k, v)
    random_program(b, rng, 140, formalism, check_every=35, classical=True)


def test_classical_only_flush_and_outcome_log():
    """Classical instructions with no quantum op around them run as their own tiny kernel; the outcome log
    keeps every measurement of every shot in draw order."""
    from dqc_simulator_b200 import timeline as tl

    b = Both("dm", 3, 6, 21)
    b.g.set_option("outcome_log", 16)
    for a in range(3):
        b.alloc_axis(a)
        b.apply_1q(a, H)
    b.measure(0, 0, -1, False, 0.0)
    b.measure(1, 1, -1, True, 0.02)
    b.check("measured")
    b.classical_op(tl.CL_CONST, 3, imm=1000.0)
    b.classical_op(tl.CL_ADDI, 4, 3, imm=135e3)
    b.classical_op(tl.CL_SEL, 5, 4, 3, cbit=0)
    b.classical_op(tl.CL_BANDN, 9, 0, 1)
    b.regs_written = {3, 4, 5}
    b.check("classical only")
    got = b.g.read_time_reg(5)
    bit0 = b.g.read_creg(0)
    assert np.array_equal(got, np.where(bit0 == 1, 136e3, 1000.0))
    log_g, log_o = b.g.read_outcomes(), b.o.read_outcomes()
    assert log_g.shape == (2, 6) and np.array_equal(log_g, log_o)
    assert np.array_equal(log_g[0], bit0)


def test_reinit_draws_fresh_shots_and_keeps_parity():
    """consecutive batches on one handle must not replay the same random shots.
    dqe_reinit advances the global shot ids by batch * shot_stride; dqe_seed replays."""
    def batch_words(e):
        for a in range(3):
            e.alloc_axis(a)
            e.apply_1q(a, H)
            e.measure(a, a, -1, False, 0.0)
        w = e.read_creg_words() if hasattr(e, "read_creg_words") else e.creg
        return w.copy(), e.full_state()

    g = _engine("dm", 3, 64, seed=7, shot_offset=0)
    g.set_option("shot_stride", 2)
    w0, _ = batch_words(g)
    g.reinit()
    w1, s1 = batch_words(g)
    assert not np.array_equal(w0, w1)
    o = OracleEngine("dm", 3, 64, seed=7, shot_offset=128)      # second batch of rank 0 of 2
    wo, so = batch_words(o)
    assert np.array_equal(w1, wo) and np.abs(s1 - so).max() < TOL
    g.seed(7, 0)
    g.reinit()                                                      # reinit after seed(…, 0): offset 128 again
    g.seed(7, 0)
    w2, _ = batch_words(g)
    assert np.array_equal(w2, w0)


@pytest.mark.parametrize("formalism,n,batch,tile", [("ket", 12, 2, 11), ("ket", 10, 3, 8), ("ket", 14, 1, 10), ("ket", 5, 4, 11),
                                                     ("dm", 6, 2, 10), ("dm", 4, 3, 8)])
def test_dense_k_qubit_blocks(formalism, n, batch, tile):
    """dqe_enqueue_kq: random dense 3- and 4-qubit matrices (unitary and not), on bits inside and outside the
    contiguous run, against the oracle -- the FP64 tensor-core path (OP_KBLOCK, SASS DMMA), including registers
    too small for one MMA tile (scalar fallback) and the two one-sided blocks of a density matrix."""
    rng = np.random.default_rng(31 + n + tile)
    b = Both(formalism, n, batch, 17, tile_bits=tile)
    for a in range(n):
        b.alloc_axis(a)
        b.apply_1q(a, rand_unitary(rng, 2))
    for rep in range(10):
        k = int(rng.choice([3, 4]))
        axes = [int(x) for x in rng.choice(n, size=k, replace=False)]
        m = rand_unitary(rng, 2 ** k)
        if rep % 4 == 3:      # a block need not be unitary (make_state_gen_op, qlib/gates.py:45-48, is not either)
            m = m @ np.diag(rng.uniform(0.5, 1.0, size=2 ** k))
        b.apply_kq(axes, m)
        if rep % 3 == 0:
            b.apply_2q(*[int(x) for x in rng.choice(n, size=2, replace=False)], CNOT)
        b.check(f"block {rep} on {axes}")


@pytest.mark.parametrize("n,tile", [(12, 11), (14, 10), (11, 8)])
def test_planner_fuses_gate_runs_into_dense_blocks(n, tile):
    """Runs of ket gates that stay inside <= 4 tile bits are composed on the host and run as one dense block;
    same state as the unfused execution and as the oracle.  Includes the non-unitary state-generation
    operator of qlib/gates.py:45-48 (diag(alpha, beta) . [[1, 1], [1, -1]])."""
    from dqc_simulator_b200.instructions import state_gen_matrix

    rng = np.random.default_rng(7 + n)

    def program(b):
        for a in range(n):
            b.alloc_axis(a)
        for layer in range(4):
            win = [int(x) for x in rng.choice(n, size=4, replace=False)]
            for _ in range(14):
                c = rng.integers(4)
                if c == 0:
                    b.apply_1q(int(rng.choice(win)), rand_unitary(rng, 2))
                elif c == 1:
                    b.apply_2q(*[int(x) for x in rng.choice(win, size=2, replace=False)], rand_unitary(rng, 4))
                elif c == 2:
                    b.apply_ctrl_1q(*[int(x) for x in rng.choice(win, size=2, replace=False)], rand_unitary(rng, 2))
                else:
                    b.apply_1q(int(rng.choice(win)), state_gen_matrix(0.8, 0.6))
            b.apply_2q(*[int(x) for x in rng.choice(n, size=2, replace=False)], CNOT)
        b.check("fused")

    state = np.random.default_rng(7 + n).bit_generator.state
    b = Both("ket", n, 2, 3, tile_bits=tile)
    program(b)
    fused = b.g.full_state()
    assert b.g.stats()["micro_ops"] < 40
    rng.bit_generator.state = state
    b2 = Both("ket", n, 2, 3, tile_bits=tile)
    b2.g.set_option("kblock", 0)
    program(b2)
    assert b2.g.stats()["micro_ops"] > b.g.stats()["micro_ops"]
    assert np.abs(fused - b2.g.full_state()).max() < TOL


@pytest.mark.parametrize("n,batch,tile,threads", [(14, 2, 11, 128), (13, 1, 10, 64), (12, 3, 8, 64), (9, 2, 11, 128),
                                                  (16, 1, 11, 256)])
def test_controlled_phase_stars(n, batch, tile, threads):
    """OP_LADDER: the controlled-phase ladders of QFT-like circuits (the reference lowers cp to
    rz . CX . rz . CX . rz, qlib/macros4parsing.py:386-463) are collected at flush into one star op per centre
    qubit, their one-bit phases fold into the neighbouring H.  Centre and leaves on thread-id bits, element
    bits and outside the tile; predicated phases; three-qubit diagonals that are no quadratic form (CCZ) and
    non-unitary diagonals stay tables.  Same state as the oracle and as the engine with the pass disabled."""
    rng = np.random.default_rng(5 + n + tile)

    def rz(t):
        return np.diag([np.exp(-0.5j * t), np.exp(0.5j * t)])

    def program(b):
        for a in range(n):
            b.alloc_axis(a)
            b.apply_1q(a, H)
        order = [int(x) for x in rng.permutation(n)]
        for stage, q in enumerate(order):
            b.apply_1q(q, H)
            for c in order[stage + 1:]:
                t = float(rng.uniform(0.1, 3.0))
                kind = rng.integers(5)
                if kind == 0:      # the reference's lowering, target q
                    b.apply_1q(q, rz(t / 2)); b.apply_2q(c, q, CNOT); b.apply_1q(q, rz(t / 2)); b.apply_2q(c, q, CNOT)
                    b.apply_1q(q, rz(t / 2))
                elif kind == 1:
                    b.apply_diag([q, c], [1, 1, 1, np.exp(1j * t)])
                elif kind == 2:
                    b.apply_ctrl_1q(c, q, np.diag([1, np.exp(1j * t)]))
                elif kind == 3:
                    b.apply_diag([c, q], np.exp(1j * rng.uniform(0, 6, size=4)))
                else:
                    b.apply_diag([q], [np.exp(0.3j * t), np.exp(-0.1j * t)])
            if stage % 4 == 1 and stage + 3 < n:
                b.apply_diag(order[stage + 1:stage + 4], [1, 1, 1, 1, 1, 1, 1, -1])              # CCZ
                b.apply_diag([q, order[stage + 1]], [1.0, 0.9, 0.8, 0.7j])                         # not unitary
            if stage == 2:
                b.apply_1q(order[0], H)       # (the phases alone leave it in |0>: outcome 1 would be impossible)
                b.measure(order[0], 3, 1, False, 0.0)
                b.set_predicate(3)
                for c in order[4:9]:
                    b.apply_diag([order[1], c], [1, 1, 1, np.exp(0.4j)])
                b.set_predicate(-1)
        b.check("stars")

    state = rng.bit_generator.state
    b = Both("ket", n, batch, 3, tile_bits=tile)
    b.g.set_option("threads", threads)
    program(b)
    with_stars = b.g.full_state()
    rng.bit_generator.state = state
    # the same program without the stars / without the block sweeps / with leftover one-bit phases sharing tables
    for opt, val in (("ladder", 0), ("kgroup", 0), ("phase_tables", 1)):
        rng.bit_generator.state = state
        b2 = Both("ket", n, batch, 3, tile_bits=tile)
        b2.g.set_option("threads", threads)
        b2.g.set_option(opt, val)
        program(b2)
        assert np.abs(with_stars - b2.g.full_state()).max() < TOL


@pytest.mark.parametrize("stream_ops", [4, 16])
def test_streaming_flush_keeps_parity(stream_ops):
    """Streaming flush (big kets: closed passes are launched while the host still enqueues; option stream_ops):
    forced on for a small register -- same amplitudes, classical bits and live axes as the oracle after every
    few steps of a random program with measurements, resets and predicated gates."""
    rng = np.random.default_rng(40 + stream_ops)
    b = Both("ket", 9, 3, 21, tile_bits=6)
    b.g.set_option("stream_min_bits", 0)
    b.g.set_option("stream_ops", stream_ops)
    random_program(b, rng, 160, "ket", check_every=20)
    b.check("streaming")


@pytest.mark.parametrize("formalism", ["ket", "dm"])
def test_forced_outcomes_bit_exact(formalism):
    n = 5
    b = Both(formalism, n, 4, 3)
    for a in range(n):
        b.alloc_axis(a)
        b.apply_1q(a, H)
    for i, (a, f) in enumerate([(0, 1), (3, 0), (1, 1)]):
        b.measure(a, i, f, False, 0.0)
    b.check()
    words = b.g.read_creg_words()
    assert np.all(words == 0b101)
    assert np.array_equal(b.g.histogram([0, 1, 2]), b.o.histogram([0, 1, 2]))
    assert np.array_equal(b.g.read_creg(0), b.o.read_creg(0))


@pytest.mark.parametrize("formalism", ["ket", "dm"])
def test_readout_matches_oracle(formalism):
    rng = np.random.default_rng(5)
    n = 6
    b = Both(formalism, n, 3, 9)
    for a in (0, 2, 3, 5, 1):
        b.alloc_axis(a)
        b.apply_1q(a, rand_unitary(rng, 2))
    for _ in range(12):
        a0, a1 = [int(x) for x in rng.choice(b.o.live, size=2, replace=False)]
        b.apply_2q(a0, a1, rand_unitary(rng, 4))
        b.depolarize([a0, a1], 0.02) if formalism == "dm" else None
    b.measure(2, 0, -1, False, 0.0)
    for axes in ([3], [5, 0], [1, 3, 0], [0, 1, 2, 3, 5]):
        for shot in range(3):
            assert np.abs(b.g.reduced_dm(axes, shot) - b.o.reduced_dm(axes, shot)).max() < TOL
        ket = rng.normal(size=2 ** len(axes)) + 1j * rng.normal(size=2 ** len(axes))
        ket /= np.linalg.norm(ket)
        fg = b.g.fidelity_ket_all(axes, ket)
        fo = [b.o.fidelity_ket(axes, ket, s) for s in range(3)]
        assert np.abs(fg - fo).max() < TOL


def test_sampling_is_sharding_invariant():
    """Shots [0, 8) in one engine == shots [0, 4) + [4, 8) in two engines (Philox keyed by
    global shot id): the property multi-GPU shot sharding relies on."""
    def run(batch, offset):
        from dqc_simulator_b200.engine import Engine

        e = Engine("ket", 3, batch, seed=42, shot_offset=offset)
        for a in range(3):
            e.alloc_axis(a)
            e.apply_1q(a, H)
        e.apply_2q(0, 1, CNOT)
        e.depolarize([0, 1], 0.3)
        for a in range(3):
            e.measure(a, a, -1, False, 0.1)
        return e.read_creg_words(), e.full_state()

    w, s = run(8, 0)
    w0, s0 = run(4, 0)
    w1, s1 = run(4, 4)
    assert np.array_equal(w, np.concatenate([w0, w1]))
    assert np.array_equal(s, np.concatenate([s0, s1]))


def test_chi_square_histogram():
    """Sampled outcome histogram vs exact probabilities of the oracle state (chi-square, 1e5 shots)."""
    from dqc_simulator_b200.engine import Engine

    rng = np.random.default_rng(0)
    n, shots = 4, 100000
    us = [rand_unitary(rng, 2) for _ in range(n)]
    u2 = rand_unitary(rng, 4)
    e = Engine("ket", n, shots, seed=2024)
    o = OracleEngine("ket", n, 1, seed=0)
    for be in (e, o):
        for a in range(n):
            be.alloc_axis(a)
            be.apply_1q(a, us[a])
        be.apply_2q(0, 2, u2)
        be.apply_2q(1, 3, CNOT)
    # exact distribution of (q3 q2 q1 q0) read as bits [0..3] -> histogram index with bit 0 = MSB
    psi = o.full_state(0).reshape(-1)
    probs = np.abs(psi) ** 2
    for a in range(n):
        e.measure(a, a, -1, False, 0.0)
    hist = e.histogram(list(range(n)))
    # histogram index = sum creg_bit[a] << (n-1-a); state index = sum bit_a << a
    exp = np.zeros(2 ** n)
    for idx in range(2 ** n):
        h = 0
        for a in range(n):
            h |= ((idx >> a) & 1) << (n - 1 - a)
        exp[h] = probs[idx] * shots
    chi2 = float(((hist - exp) ** 2 / exp).sum())
    # dof = 15; P(chi2 > 50) ~ 1e-5
    assert hist.sum() == shots
    assert chi2 < 50.0, chi2


def test_errors_are_reported_not_thrown_across_abi():
    from dqc_simulator_b200.engine import Engine, EngineError

    e = Engine("ket", 3, 1)
    with pytest.raises(EngineError, match="holds no qubit"):
        e.apply_1q(0, X)
    e.alloc_axis(0)
    with pytest.raises(EngineError, match="already live"):
        e.alloc_axis(0)
    with pytest.raises(EngineError, match="dm-only"):
        e.kraus_1q(0, [np.eye(2)])
    e.apply_1q(0, X)
    assert abs(e.full_state(0)[0, 1] - 1) < 1e-15


@pytest.mark.parametrize("ctrl", [-1, 2])
def test_peer_memory_gate_on_a_global_axis(ctrl):
    """k_peer_1q (fused gate + transfer for sharded state vectors): two engines on one GPU play
    ranks 0 and 1 of a register whose top axis is the rank bit; each 'rank' updates both shards on
    its half of the index range.  Result = the gate applied to axis m of an (m+1)-qubit oracle ket."""
    import ctypes

    from dqc_simulator_b200.engine import Engine

    m = 6
    rng = np.random.default_rng(4)
    psi = rng.normal(size=2 ** (m + 1)) + 1j * rng.normal(size=2 ** (m + 1))
    psi /= np.linalg.norm(psi)
    u = rand_unitary(rng, 2)
    shards = []
    for r in range(2):
        e = Engine("ket", m, 1)
        e.set_state(0, psi[r * 2 ** m:(r + 1) * 2 ** m], list(range(m)))
        shards.append(e)
    ptrs = []
    for e in shards:
        dev, nbytes = ctypes.c_void_p(), ctypes.c_int64()
        e._ck(e._lib.dqe_state_ptr(e._h, ctypes.byref(dev), ctypes.byref(nbytes)))
        ptrs.append(dev.value)
    for r in range(2):
        shards[r].peer_1q(ptrs[1 - r], my_bit=r, half=r, m=u, ctrl_axis=ctrl)
    for e in shards:
        e.sync()
    got = np.concatenate([e.full_state(0)[0] for e in shards])
    want = psi.reshape(2, -1).copy()
    if ctrl < 0:
        want = u @ want
    else:
        sel = ((np.arange(2 ** m) >> ctrl) & 1).astype(bool)
        want[:, sel] = u @ want[:, sel]
    assert np.abs(got - want.reshape(-1)).max() < 1e-12


def test_chi_square_dm_trajectory_outcomes():
    """Density-matrix trajectories: the joint histogram of the four cat-comm measurement outcomes of the noisy
    GHZ-5 / 3-QPU circuit (16 outcome strings, read from the engine's outcome log) against the histogram of an
    INDEPENDENT sample of the oracle (other seed, 6000 shots): two-sample chi-square over the 16 strings, and the
    same statistic against the engine's own second half as the yardstick of what "same distribution" gives."""
    from dqc_simulator_b200 import executor, experiment
    from dqc_simulator_b200.engine import Engine

    ops, _ = experiment.load_c2("ghz")
    dqc = experiment.c2_hardware(noisy=True, F_werner=0.9, meas_error_prob=0.05)

    def strings(be):
        ex = executor.DQCExecutor(ops, dqc, be).run()
        assert ex.measurements == 4
        out = np.asarray(be.read_outcomes(), dtype=np.int64)         # (4, batch)
        return (out[0] << 3) | (out[1] << 2) | (out[2] << 1) | out[3]

    def chi2(a, b):
        n1, n2 = a.sum(), b.sum()
        return float((((np.sqrt(n2 / n1) * a - np.sqrt(n1 / n2) * b) ** 2) / np.maximum(a + b, 1.0)).sum())

    e = Engine("dm", 7, 40000, seed=99)
    e.set_option("outcome_log", 8)
    s_e = strings(e)
    a = np.bincount(s_e, minlength=16).astype(float)
    b = np.zeros(16)
    for chunk in range(12):
        o = OracleEngine("dm", 7, 500, seed=7, shot_offset=chunk * 500)
        b += np.bincount(strings(o), minlength=16)
    assert chi2(a, b) < 45.0, (a, b)          # dof = 15, P(chi2 > 45) ~ 8e-5
    half = chi2(np.bincount(s_e[:20000], minlength=16).astype(float), np.bincount(s_e[20000:], minlength=16).astype(float))
    assert half < 45.0
    # and the same seed reproduces the oracle's strings shot for shot (the sampling path itself, bit-exact)
    o = OracleEngine("dm", 7, 256, seed=99)
    assert np.array_equal(strings(o), s_e[:256])


@pytest.mark.parametrize("L", [0, 3, 5])
def test_peer_memory_axis_swap(L):
    """k_peer_swap: exchange the rank-bit axis with local axis L in place; two engines on one GPU play
    the two ranks, each handling half of the pairs."""
    import ctypes

    from dqc_simulator_b200.engine import Engine

    m = 6
    rng = np.random.default_rng(8)
    psi = rng.normal(size=2 ** (m + 1)) + 1j * rng.normal(size=2 ** (m + 1))
    shards, ptrs = [], []
    for r in range(2):
        e = Engine("ket", m, 1)
        e.set_state(0, psi[r * 2 ** m:(r + 1) * 2 ** m], list(range(m)))
        dev, nbytes = ctypes.c_void_p(), ctypes.c_int64()
        e._ck(e._lib.dqe_state_ptr(e._h, ctypes.byref(dev), ctypes.byref(nbytes)))
        shards.append(e)
        ptrs.append(dev.value)
    for r in range(2):
        shards[r].peer_swap(ptrs[1 - r], my_bit=r, half=r, local_axis=L)
    for e in shards:
        e.sync()
    got = np.concatenate([e.full_state(0)[0] for e in shards])
    idx = np.arange(2 ** (m + 1))
    g, l = (idx >> m) & 1, (idx >> L) & 1
    src = (idx & ~((1 << m) | (1 << L))) | (l << m) | (g << L)     # swap bits m and L
    assert np.array_equal(got, psi[src])


def test_replayed_plan_equals_a_fresh_walk():
    """experiment.run_shot_batch keeps the plan of the first batch and replays it on the next block of shot ids
    (dqe_replay): batch by batch the same classical registers and fidelities as an engine that walks and plans the
    program again every time."""
    from dqc_simulator_b200 import experiment

    ops, _ = experiment.load_c2("qft")
    dqc = experiment.c2_hardware(noisy=True)
    data, ideal = experiment.ideal_output_ket(ops, lambda f, n, b: _engine(f, n, b))
    a = _engine("dm", 7, 64, seed=5, shot_offset=0)
    b = _engine("dm", 7, 64, seed=5, shot_offset=0)
    for batch in range(3):
        ra = experiment.run_shot_batch(ops, dqc, a, data, ideal)                      # batches 1, 2: replayed
        rb = experiment.run_shot_batch(ops, dqc, b, data, ideal, reuse_plan=False)    # walked every time
        assert np.array_equal(ra.creg, rb.creg), batch
        assert np.array_equal(ra.fidelity, rb.fidelity), batch
        assert sorted(a.live) == sorted(b.live)
    assert a.stats()["api_ops"] == b.stats()["api_ops"]
    with pytest.raises(Exception, match="fresh"):
        a.replay()                                   # not reinitialised
    a.set_option("cxperm", 1)                        # any option change drops the plan
    a.reinit()
    with pytest.raises(Exception, match="no cached plan"):
        a.replay()


@pytest.mark.parametrize("circuit,tma", [("grover", 1), ("qft", 1), ("qft", 0)])
def test_chained_passes_equal_separate_launches(circuit, tma):
    """Pass chains (option chain; off by default, it measured slower: DESIGN.md section 9): consecutive passes that keep one tile per shot in the same layout run as
    ONE launch -- the tile stays in shared memory, only the micro-program is re-staged per link.  Same arithmetic in
    the same order: states, classical registers and fidelities are bit-identical to one launch per pass, on a fresh
    walk and on a replayed plan, with far fewer launches."""
    from dqc_simulator_b200 import experiment

    ops, _ = experiment.load_c2(circuit)
    dqc = experiment.c2_hardware(noisy=True)
    data, ideal = experiment.ideal_output_ket(ops, lambda f, n, b: _engine(f, n, b))
    a = _engine("dm", 7, 48, seed=11, shot_offset=0)
    b = _engine("dm", 7, 48, seed=11, shot_offset=0)
    a.set_option("chain", 1)
    for e in (a, b):
        e.set_option("tma", tma)
    for batch in range(2):                               # batch 1 replays the cached plan
        ra = experiment.run_shot_batch(ops, dqc, a, data, ideal)
        rb = experiment.run_shot_batch(ops, dqc, b, data, ideal)
        assert np.array_equal(ra.creg, rb.creg), batch
        assert np.array_equal(ra.fidelity, rb.fidelity), batch
        assert np.array_equal(a.full_state(), b.full_state()), batch
    sa, sb = a.stats(), b.stats()
    assert sa["passes"] == sb["passes"] and sa["micro_ops"] == sb["micro_ops"] and sa["amp_updates"] == sb["amp_updates"]
    assert sa["kernel_launches"] < sb["kernel_launches"] - sb["passes"] // 2


def test_recorded_program_replays_bit_identically():
    """Engine.record / play on the device: the second batch of a sampled, noisy QFT-5 program replayed from the
    recorded C-ABI stream equals the second batch of an engine that walks the program again (same global shot ids)."""
    from dqc_simulator_b200 import executor, experiment

    ops, _ = experiment.load_c2("qft")
    dqc = experiment.c2_hardware(noisy=True)
    a = _engine("dm", 7, 24, seed=3, shot_offset=0)
    b = _engine("dm", 7, 24, seed=3, shot_offset=0)
    with a.record() as rec:
        executor.DQCExecutor(ops, dqc, a).run()
        a.flush()
    executor.DQCExecutor(ops, dqc, b).run()
    assert rec.valid
    a.reinit(); b.reinit()
    a.play(rec)
    executor.DQCExecutor(ops, dqc, b).run()
    assert np.array_equal(a.read_creg_words(), b.read_creg_words())
    assert np.array_equal(a.full_state(), b.full_state())
    assert sorted(a.live) == sorted(b.live)


@pytest.mark.parametrize("formalism,n,batch,tile", [("dm", 6, 3, 10), ("dm", 7, 2, 8), ("dm", 5, 4, 12), ("dm", 9, 1, 11),
                                                     ("ket", 10, 6, 8), ("ket", 13, 3, 11)])
def test_two_qubit_pauli_channel(formalism, n, batch, tile):
    """dqe_enqueue_pauli_channel_k (k, axes, 4^k weights): correlated two-qubit Pauli noise between gates
    and measurements -- dm: one dense 16 x 16 superoperator block on (row a, row b, col a, col b), ket: one Pauli
    pair drawn per shot -- against the oracle; the product case equals two one-qubit channels, the uniform case the
    two-qubit depolarising channel of noise_models.py:295-366."""
    rng = np.random.default_rng(77 + n + tile)
    b = Both(formalism, n, batch, 19, tile_bits=tile)
    for a in range(n):
        b.alloc_axis(a)
        b.apply_1q(a, rand_unitary(rng, 2))
    for step in range(30):
        a0, a1 = [int(x) for x in rng.choice(n, size=2, replace=False)]
        kind = step % 5
        if kind == 0:
            b.apply_2q(a0, a1, rand_unitary(rng, 4))
        elif kind == 1:
            b.pauli_channel([a0, a1], rng.dirichlet([12] + [1] * 15))
        elif kind == 2:
            b.apply_2q(a0, a1, CNOT)
            w = np.zeros(16); w[0], w[5], w[10], w[15] = 0.7, 0.1, 0.1, 0.1        # correlated: XX, YY, ZZ
            b.pauli_channel([a1, a0], w)
        elif kind == 3:
            b.apply_1q(a0, rand_unitary(rng, 2))
            b.pauli_channel([a0], rng.dirichlet([8, 1, 1, 1]))                      # k = 1 through the same entry point
        else:
            b.measure(a0, step % 30, -1 if step % 2 else int(rng.integers(2)), False, 0.01)
        if step % 10 == 9:
            b.check(f"step {step}")
    b.check("end")
    if formalism == "dm":
        # product weights = two one-qubit channels; uniform non-identity weights = two-qubit depolarising
        e1, e2 = _engine("dm", 4, 1, seed=1), _engine("dm", 4, 1, seed=1)
        for e in (e1, e2):
            for a in range(4):
                e.alloc_axis(a)
                e.apply_1q(a, rand_unitary(np.random.default_rng(a), 2))
            e.apply_2q(0, 1, CNOT); e.apply_2q(2, 3, CNOT); e.apply_2q(1, 2, rand_unitary(np.random.default_rng(9), 4))
        pa, pb = rng.dirichlet([8, 1, 1, 1]), rng.dirichlet([6, 2, 1, 1])
        e1.pauli_channel([1, 3], np.outer(pa, pb).reshape(16))
        e2.pauli_channel(1, pa); e2.pauli_channel(3, pb)
        p = 0.23
        w = np.full(16, p / 16); w[0] = 1 - 15 * p / 16
        e1.pauli_channel([0, 2], w)
        e2.depolarize([0, 2], p)
        assert np.abs(e1.full_state() - e2.full_state()).max() < TOL
