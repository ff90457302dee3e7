"""NumPy complex128 oracle for the dqc_simulator state-evolution hot path.

TEST INFRASTRUCTURE ONLY -- this is the checker, never the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs
may import it.  The product package (``dqc_simulator_b200``) never does and fails loudly
when its CUDA library is missing.

What it restates
----------------
The arithmetic of the reference's hot path lives in the closed ``netsquid==1.1.7`` wheel
(``/root/reference/uv.lock:39-51``), which cannot be installed here (cp39-only, login-gated
index).  This file therefore restates the *published* semantics of the
NetSquid calls the reference makes, anchored on the reference's own call sites:

* gate application ``qubitapi.operate`` -- reached from ``QuantumProgram.apply`` at
  ``software/dqc_control.py:191-196, 295-298, 348, 365-366, 390, 442-459, 473, 498-503, 598``;
  operator matrices from ``qlib/gates.py:29-220``; big-endian qubit order
  (first listed qubit = most significant kron factor; pinned by
  ``tests/unit_tests/qlib/test_circuit_identities.py:93-96``).
* ``Operator.ctrl`` = |0><0| (x) I + |1><1| (x) U with control first (``qlib/gates.py:96-98,151``).
* Z-basis ``INSTR_MEASURE`` with collapse and optional discard
  (``hardware/quantum_processors.py:414-422``: ``discard=True``, noise applied *before*).
* k-qubit depolarising ``NDimDepolarNoiseModel.error_operation``
  (``hardware/noise_models.py:343-366``): uniform mixture over the 4^k Pauli strings with
  total error weight p, i.e. (1-p) rho + p Tr_S(rho) (x) I/2^k -- the same closed form the
  deprecated ``apply_analytical_depolarisation2dm`` computes (``noise_models.py:33-131``).
* NetSquid ``DepolarNoiseModel`` (1-qubit gate and memory noise,
  ``quantum_processors.py:868-877``): (1-p) rho + p I/2, with p = 1 - exp(-dt*1e-9*rate)
  for memory noise (``noise_models.py:248-253``).
* ``MeasurementNoiseModel`` (``noise_models.py:295-307``): Pauli channel (1-e, e, 0, 0).
* 2-qubit state injection from ``BlackBoxEntanglingQsourceConnection``
  (``hardware/connections.py:145-150``) of ``ks.b00`` or ``werner_state(F)``
  (``qlib/states.py:63-84``).
* read-out ``qapi.reduced_dm`` / ``qapi.fidelity(..., squared=True)``
  (``util/helper.py:151, 209-214``; ``examples/run_experiment.py:74-75``).

DM formalism applies every channel analytically; KET formalism samples one branch per shot
(SURVEY.md Appendix A).  Random draws come from ``oracle/philox.py`` with the same
(seed, global shot id, draw index) keying as the CUDA engine.

Parity pinning
--------------
Pinned against every closed-form state the reference's own tests assert for this path
(``tests/test_oracle_known_answers.py`` lifts them with file:line citations) and the
circuit-level states of ``test_dqc_control.py`` / ``test_circuit_identities.py``, and
against real NetSquid output: the 16-digit noisy end-to-end fidelities the reference's
examples print (``examples/getting_started.py:128-139``, ``from_monolithic_circuit.py:143-155``,
``run_experiment.py:80-93`` GHZ-5) are reproduced to <= 4e-16 by the one trajectory of forced
measurement outcomes that produced them (``tests/test_reference_printed_fidelities.py``) --
executor timeline, channel order, Werner ebits and this arithmetic all have to agree with
NetSquid's for that.  The reference's own tests pin states only to ~1e-5 fidelity and up to
global phase; amplitude-level (1e-10) agreement beyond those numbers is engine-vs-oracle
only.  Histogram sampling is *parity unpinned* upstream (no sampling test exists in the
reference).
"""
import numpy as np

from . import philox

KET = 0
DM = 1

_PAULI = [
    np.array([[1, 0], [0, 1]], dtype=complex),
    np.array([[0, 1], [1, 0]], dtype=complex),
    np.array([[0, -1j], [1j, 0]], dtype=complex),
    np.array([[1, 0], [0, -1]], dtype=complex),
]


def _apply_matrix(arr, mat, dims):
    """Contract a (2^k x 2^k) matrix into the listed tensor dims of ``arr`` (first dim listed
    is the most significant index of the matrix)."""
    k = len(dims)
    m = np.asarray(mat, dtype=complex).reshape((2,) * (2 * k))
    out = np.tensordot(m, arr, axes=(list(range(k, 2 * k)), list(dims)))
    return np.moveaxis(out, list(range(k)), list(dims))


class OracleEngine:
    """Batched ket / density-matrix register with dynamic live-qubit list.

    Axis ids are small integers chosen by the caller (they mirror the CUDA engine's axes);
    ``self.live`` lists the live axes in tensor order.  ket: ``state.shape == (B, 2, ..., 2)``;
    dm: ``(B, 2, ..., 2 [rows], 2, ..., 2 [cols])``.
    """

    def __init__(self, formalism, max_qubits, batch, seed=0, shot_offset=0):
        self.formalism = KET if formalism in (KET, "ket") else DM
        self.max_qubits = int(max_qubits)
        self.batch = int(batch)
        self.live = []
        self.state = np.ones((self.batch,), dtype=complex)
        self.creg = np.zeros((self.batch,), dtype=np.uint64)
        self._pred_bit = -1
        # per-shot classical register file (simulated times / channel strengths), dqe_classical_op
        self.n_time_regs = 128
        self.treg = np.zeros((self.n_time_regs, self.batch), dtype=np.float64)
        self.outcomes = []   # one int8 array (batch) per measurement, in draw order (dqe_read_outcomes)
        self.seed(seed, shot_offset)

    # ------------------------------------------------------------------ bookkeeping
    def seed(self, seed, shot_offset=0):
        self._seed = int(seed)
        self._shot_ids = np.arange(self.batch, dtype=np.int64) + int(shot_offset)
        self._draw = 0

    def _next_uniforms(self):
        u = philox.shot_uniforms(self._seed, self._shot_ids, self._draw)
        self._draw += 1
        return u

    @property
    def n_live(self):
        return len(self.live)

    def _dim(self, axis):
        """tensor dim (excluding batch) of the ket/row index of a live axis."""
        return 1 + self.live.index(axis)

    def _cdim(self, axis):
        return 1 + self.n_live + self.live.index(axis)

    def _bshape(self, v):
        return np.asarray(v).reshape((self.batch,) + (1,) * (self.state.ndim - 1))

    def alloc_axis(self, axis=-1):
        """Make an axis live in |0>.  Mirrors ``dqe_alloc_axis``."""
        if axis < 0:
            axis = min(a for a in range(self.max_qubits) if a not in self.live)
        if axis in self.live or not 0 <= axis < self.max_qubits:
            raise ValueError(f"axis {axis} unavailable")
        z = np.zeros((2,), dtype=complex)
        z[0] = 1
        k = self.n_live
        if self.formalism == KET:
            self.state = np.multiply.outer(self.state, z)
        else:
            st = np.multiply.outer(np.multiply.outer(self.state, z), z)
            # dims now (B, rows k, cols k, newrow, newcol) -> move newrow after rows
            st = np.moveaxis(st, 1 + 2 * k, 1 + k)
            self.state = st
        self.live.append(axis)
        return axis

    def free_axis(self, axis):
        """Discard a live qubit: DM partial trace; ket = unrecorded sampled measurement."""
        if self.formalism == DM:
            self.state = np.trace(self.state, axis1=self._dim(axis), axis2=self._cdim(axis))
            self.live.remove(axis)
        else:
            self._measure(axis, None, -1, True, 0.0)

    def reset(self, axis):
        """INSTR_INIT on an occupied position: replace the occupant by a fresh |0>."""
        if axis in self.live:
            self.free_axis(axis)
        self.alloc_axis(axis)

    # ------------------------------------------------------------------ unitaries
    def apply_1q(self, axis, m):
        self._apply([axis], np.asarray(m, dtype=complex).reshape(2, 2))

    def apply_2q(self, a, b, m):
        self._apply([a, b], np.asarray(m, dtype=complex).reshape(4, 4))

    def apply_ctrl_1q(self, c, t, m):
        full = np.eye(4, dtype=complex)
        full[2:, 2:] = np.asarray(m, dtype=complex).reshape(2, 2)
        self._apply([c, t], full)

    def apply_diag(self, axes, d):
        self._apply(list(axes), np.diag(np.asarray(d, dtype=complex)))

    def apply_kq(self, axes, m):
        axes = list(axes)
        self._apply(axes, np.asarray(m, dtype=complex).reshape(2 ** len(axes), 2 ** len(axes)))

    def set_predicate(self, creg_bit):
        """Ops enqueued until ``set_predicate(-1)`` act only on shots whose creg bit is 1."""
        self._pred_bit = int(creg_bit)

    def _commit(self, new, predicate=None):
        if self._pred_bit >= 0:
            pb = ((self.creg >> np.uint64(self._pred_bit)) & np.uint64(1)).astype(bool)
            predicate = pb if predicate is None else (pb & predicate)
        if predicate is None:
            self.state = new
        else:
            self.state = np.where(self._bshape(predicate), new, self.state)

    def _apply(self, axes, m, predicate=None):
        new = _apply_matrix(self.state, m, [self._dim(a) for a in axes])
        if self.formalism == DM:
            new = _apply_matrix(new, m.conj(), [self._cdim(a) for a in axes])
        self._commit(new, predicate)

    def cond_1q(self, axis, m, creg_bit):
        """Classically controlled gate: applied on the shots whose creg bit is 1
        (``dqc_control.py:347-348, 387-390, 497-503``)."""
        pred = ((self.creg >> np.uint64(creg_bit)) & np.uint64(1)).astype(bool)
        self._apply([axis], np.asarray(m, dtype=complex).reshape(2, 2), predicate=pred)

    # ------------------------------------------------------------------ channels
    def pauli_channel(self, axis, p):
        p = np.asarray(p, dtype=float)
        if isinstance(axis, (list, tuple, np.ndarray)) and len(axis) == 2:
            # two-qubit Pauli channel sum_ab p[4a + b] (P_a x P_b) rho (P_a x P_b) (restates what NetSquid's
            # qubitapi.apply_pauli_noise does qubit by qubit, noise_models.py:343-366, for correlated weights);
            # ket: one pair per shot from the shot's next uniform, cumulative order s = 4a + b
            a0, a1 = int(axis[0]), int(axis[1])
            if self.formalism == DM:
                acc = 0
                for s_ in range(16):
                    w = p[s_]
                    if w == 0.0:
                        continue
                    Pa, Pb = _PAULI[s_ >> 2], _PAULI[s_ & 3]
                    t = _apply_matrix(self.state, Pa, [self._dim(a0)])
                    t = _apply_matrix(t, Pb, [self._dim(a1)])
                    t = _apply_matrix(t, Pa.conj(), [self._cdim(a0)])
                    acc = acc + w * _apply_matrix(t, Pb.conj(), [self._cdim(a1)])
                self._commit(acc)
            else:
                u, _ = self._next_uniforms()
                cum = np.cumsum(p)
                choice = np.zeros(u.shape, dtype=int)
                for i in range(15):
                    choice = choice + (u >= cum[i])
                for q, ax in ((choice >> 2, a0), (choice & 3, a1)):
                    for idx in (1, 2, 3):
                        sel = q == idx
                        if sel.any():
                            self._apply([ax], _PAULI[idx], predicate=sel)
            return
        if isinstance(axis, (list, tuple, np.ndarray)):
            axis = int(axis[0])
        if self.formalism == DM:
            acc = 0
            for w, P in zip(p, _PAULI):
                if w != 0.0:
                    t = _apply_matrix(self.state, P, [self._dim(axis)])
                    acc = acc + w * _apply_matrix(t, P.conj(), [self._cdim(axis)])
            self._commit(acc)
        else:
            u, _ = self._next_uniforms()
            cum = np.cumsum(p)
            choice = (u >= cum[0]).astype(int) + (u >= cum[1]) + (u >= cum[2])
            for idx in (1, 2, 3):
                sel = choice == idx
                if sel.any():
                    self._apply([axis], _PAULI[idx], predicate=sel)

    def depolarize(self, axes, p):
        axes = list(axes)
        k = len(axes)
        if self.formalism == DM:
            nl = self.n_live
            rows = list(range(1, nl + 1))
            cols = list(range(nl + 1, 2 * nl + 1))
            st = self.state
            in_rows, in_cols = list(rows), list(cols)
            operands = []
            for j, a in enumerate(axes):
                pos = self.live.index(a)
                t = 2 * nl + 1 + j  # fresh label: traced over in the input
                in_rows[pos] = t
                in_cols[pos] = t
                operands += [np.eye(2, dtype=complex) / 2.0, [rows[pos], cols[pos]]]
            mixed = np.einsum(st, [0] + in_rows + in_cols, *operands, [0] + rows + cols)
            self._commit((1.0 - p) * st + p * mixed)
        else:
            # sampled Pauli string, weights of noise_models.py:353-366
            u, _ = self._next_uniforms()
            n = 4 ** k
            w = p / n
            w0 = 1.0 - (n - 1) * w
            # string index s in [0, n): 0 = identity
            s = np.where(u < w0, 0, 1 + np.floor((u - w0) / w).astype(np.int64) if w > 0 else 0)
            s = np.minimum(s, n - 1)
            for i, a in enumerate(axes):
                digit = (s // (4 ** (k - 1 - i))) % 4
                for idx in (1, 2, 3):
                    sel = digit == idx
                    if sel.any():
                        self._apply([a], _PAULI[idx], predicate=sel)

    def kraus_1q(self, axis, ks):
        ks = np.asarray(ks, dtype=complex).reshape(-1, 2, 2)
        if self.formalism != DM:
            raise NotImplementedError("generic Kraus channels are DM-only")
        acc = 0
        for K in ks:
            t = _apply_matrix(self.state, K, [self._dim(axis)])
            acc = acc + _apply_matrix(t, K.conj(), [self._cdim(axis)])
        self._commit(acc)

    # ------------------------------------------------------------------ per-shot classical machine
    def _bit(self, b):
        return ((self.creg >> np.uint64(b)) & np.uint64(1)).astype(bool)

    def _set_bit(self, b, v):
        bit = np.uint64(1) << np.uint64(b)
        self.creg = (self.creg & ~bit) | (v.astype(np.uint64) << np.uint64(b))

    def classical_op(self, op, dst, a=0, b=0, cbit=-1, imm=0.0):
        """Restates ``dqe_classical_op`` (include/dqc_engine.h): one instruction of the per-shot register
        machine that carries the simulated timelines of batched shots (dqc_simulator_b200/timeline.py);
        opcodes are timeline.CL_*.  Evaluated eagerly here, per shot, in the engine's operation order."""
        R = self.treg
        if op == 1:      # CONST
            R[dst] = imm
        elif op == 2:    # ADDI
            R[dst] = R[a] + imm
        elif op == 3:    # ADD
            R[dst] = R[a] + R[b] + imm
        elif op == 4:    # SUB
            R[dst] = R[a] - R[b] + imm
        elif op == 5:    # RSUBI
            R[dst] = imm - R[a]
        elif op == 6:    # MAX
            R[dst] = np.maximum(R[a], R[b] + imm)
        elif op == 7:    # MAXI
            R[dst] = np.maximum(R[a], imm)
        elif op == 8:    # SEL
            R[dst] = np.where(self._bit(cbit), R[a] + imm, R[b])
        elif op == 12:   # ADDIF
            R[dst] = R[a] + np.where(self._bit(cbit), imm, 0.0)
        elif op == 9:    # CEILQ: handshake re-sends, physical_layer.py:269-335
            R[dst] = np.maximum(0.0, np.ceil(R[a] / imm - 1e-12))
        elif op == 10:   # FMAI
            R[dst] = R[a] * imm + R[b]
        elif op == 11:   # QEXP: noise_models.py:248-253
            R[dst] = 1.0 - np.exp(-np.maximum(0.0, R[a]) * imm)
        elif op == 16:   # BAND
            self._set_bit(dst, self._bit(a) & self._bit(b))
        elif op == 17:   # BANDN
            self._set_bit(dst, self._bit(a) & ~self._bit(b))
        elif op == 18:   # BOR
            self._set_bit(dst, self._bit(a) | self._bit(b))
        elif op == 19:   # BNOT
            self._set_bit(dst, ~self._bit(a))
        elif op == 20:   # BCOPY
            self._set_bit(dst, self._bit(a))
        else:
            raise ValueError(f"unknown classical opcode {op}")

    def classical_qexp2(self, dst, a, b, offset, rate):
        ra = 0.0 if a == 255 else self.treg[a]
        rb = 0.0 if b == 255 else self.treg[b]
        self.treg[dst] = 1.0 - np.exp(-np.maximum(0.0, ra - rb + offset) * rate)

    def classical_fence(self):
        pass

    def classical_rand(self, kind, dst, p, kmax=0.0):
        """dqe_classical_rand: per-shot Philox uniform of the next draw; bit <- (u < p) or the geometric attempt count."""
        u, _ = self._next_uniforms()
        if kind == 0:
            bit = np.uint64(1) << np.uint64(dst)
            self.creg = (self.creg & ~bit) | np.where(u < p, bit, np.uint64(0))
        else:
            lnq = -1e300 if p >= 1.0 else (-1e-300 if p <= 0.0 else np.log1p(-p))
            with np.errstate(divide="ignore"):
                k = np.where(u > 0.0, 1.0 + np.floor(np.log(np.maximum(u, 1e-300)) / lnq), kmax)
            self.treg[dst] = np.minimum(kmax, k)

    def read_time_reg(self, reg):
        return self.treg[int(reg)].copy()

    def dyn_channel(self, axis, kind, reg):
        """One-qubit channel whose strength q is read per shot from classical register ``reg``:
        kind 0 depolarising (1-q) rho + q I/2, 1 dephasing (off-diagonals x (1-q)), 2 amplitude damping
        (gamma = q).  DM: analytic; ket: one branch sampled per shot (depolarising / dephasing only)."""
        q = self.treg[int(reg)]
        if self.formalism == DM:
            st = np.moveaxis(self.state, [self._dim(axis), self._cdim(axis)], [1, 2])
            new = st.copy()
            qb = q.reshape((-1,) + (1,) * (st.ndim - 3))
            e00, e01, e10, e11 = st[:, 0, 0], st[:, 0, 1], st[:, 1, 0], st[:, 1, 1]
            if kind == 0:
                new[:, 0, 0] = (1 - qb / 2) * e00 + (qb / 2) * e11
                new[:, 1, 1] = (1 - qb / 2) * e11 + (qb / 2) * e00
                new[:, 0, 1] = (1 - qb) * e01
                new[:, 1, 0] = (1 - qb) * e10
            elif kind == 1:
                new[:, 0, 1] = (1 - qb) * e01
                new[:, 1, 0] = (1 - qb) * e10
            elif kind == 3:
                new[:, 0, 1] = (1 - 2 * qb) * e01
                new[:, 1, 0] = (1 - 2 * qb) * e10
            elif kind == 2:
                new[:, 0, 0] = e00 + qb * e11
                new[:, 1, 1] = (1 - qb) * e11
                new[:, 0, 1] = np.sqrt(1 - qb) * e01
                new[:, 1, 0] = np.sqrt(1 - qb) * e10
            else:
                raise ValueError("dyn_channel kind")
            self._commit(np.moveaxis(new, [1, 2], [self._dim(axis), self._cdim(axis)]))
            return
        u, _ = self._next_uniforms()
        if kind == 0:
            w = q / 4.0
            w0 = 1.0 - 3.0 * w
            with np.errstate(divide="ignore", invalid="ignore"):
                s = np.where(u < w0, 0, np.where(w > 0, 1 + np.floor((u - w0) / w), 0)).astype(np.int64)
            s = np.minimum(s, 3)
        elif kind == 1:
            s = np.where(u >= 1.0 - q / 2.0, 3, 0)
        elif kind == 3:
            s = np.where(u >= 1.0 - q, 3, 0)
        else:
            raise NotImplementedError("amplitude damping is DM-only")
        for idx in (1, 2, 3):
            sel = s == idx
            if sel.any():
                self._apply([axis], _PAULI[idx], predicate=sel)

    # ------------------------------------------------------------------ injection
    def inject_2q(self, a, b, state, is_dm):
        """Place a 2-qubit state on two free axes (a = most significant)."""
        for x in (a, b):
            if x in self.live:
                raise ValueError(f"axis {x} is occupied")
        k = self.n_live
        if self.formalism == DM:
            sig = np.asarray(state, dtype=complex)
            if not is_dm:
                v = sig.reshape(4)
                sig = np.outer(v, v.conj())
            sig = sig.reshape(2, 2, 2, 2)  # ra rb ca cb
            st = np.multiply.outer(self.state, sig)
            nd = st.ndim
            st = np.moveaxis(st, [nd - 4, nd - 3], [1 + k, 2 + k])  # rows after rows
            self.state = st
        else:
            if is_dm:
                sig = np.asarray(state, dtype=complex).reshape(4, 4)
                lam, vec = np.linalg.eigh(sig)
                lam = np.clip(lam[::-1], 0, None)
                vec = vec[:, ::-1]
                u, _ = self._next_uniforms()
                cum = np.cumsum(lam / lam.sum())
                choice = (u >= cum[0]).astype(int) + (u >= cum[1]) + (u >= cum[2])
                kets = vec.T[choice]  # (B, 4)
                st = self.state[..., None] * kets.reshape((self.batch,) + (1,) * k + (4,))
                self.state = st.reshape(st.shape[:-1] + (2, 2))
            else:
                v = np.asarray(state, dtype=complex).reshape(2, 2)
                self.state = np.multiply.outer(self.state, v)
        self.live += [a, b]

    # ------------------------------------------------------------------ measurement
    def measure(self, axis, creg_bit, forced=-1, discard=False, flip_prob=0.0):
        return self._measure(axis, creg_bit, forced, discard, flip_prob)

    def prob_one(self, axis):
        if self.formalism == KET:
            p = np.abs(self.state) ** 2
            d = self._dim(axis)
            other = tuple(i for i in range(1, p.ndim) if i != d)
            pr = p.sum(axis=other)  # (B, 2)
        else:
            k = self.n_live
            diag = np.einsum(
                self.state, [0] + list(range(1, k + 1)) + list(range(1, k + 1)),
                [0] + list(range(1, k + 1))).real
            d = self._dim(axis)
            other = tuple(i for i in range(1, diag.ndim) if i != d)
            pr = diag.sum(axis=other)
        return pr[:, 0], pr[:, 1]

    def _measure(self, axis, creg_bit, forced, discard, flip_prob):
        u, u2 = self._next_uniforms()
        if self.formalism == KET and flip_prob > 0.0:
            self._apply([axis], _PAULI[1], predicate=(u2 < flip_prob))
        p0, p1 = self.prob_one(axis)
        e = flip_prob if self.formalism == DM else 0.0
        q0 = (1 - e) * p0 + e * p1
        q1 = (1 - e) * p1 + e * p0
        tot = q0 + q1
        if forced is None or forced < 0:
            m = (u * tot < q1).astype(np.int64)
        else:
            m = np.full((self.batch,), int(forced), dtype=np.int64)
        pm = np.where(m == 1, q1, q0)
        d = self._dim(axis)
        st = np.moveaxis(self.state, d, 1)
        if self.formalism == KET:
            sel = np.take_along_axis(st, m.reshape((-1, 1) + (1,) * (st.ndim - 2)), axis=1)[:, 0]
            with np.errstate(divide="ignore", invalid="ignore"):
                sc = np.where(pm > 0, 1.0 / np.sqrt(pm), 0.0)
            kept = sel * sc.reshape((-1,) + (1,) * (sel.ndim - 1))
        else:
            cd = self._cdim(axis)
            st = np.moveaxis(st, cd, 2)  # (B, r, c, ...)
            b00 = st[:, 0, 0]
            b11 = st[:, 1, 1]
            mm = m.reshape((-1,) + (1,) * (b00.ndim - 1)).astype(bool)
            same = np.where(mm, b11, b00)
            flip = np.where(mm, b00, b11)
            with np.errstate(divide="ignore", invalid="ignore"):
                sc = np.where(pm > 0, 1.0 / pm, 0.0)
            kept = ((1 - e) * same + e * flip) * sc.reshape((-1,) + (1,) * (same.ndim - 1))
        self.live.remove(axis)
        self.state = kept
        if not discard:
            # re-attach the qubit in |m>
            self.alloc_axis(axis)
            self._apply([axis], _PAULI[1], predicate=(m == 1))
        if creg_bit is not None and creg_bit >= 0:
            bit = np.uint64(1) << np.uint64(creg_bit)
            self.creg = (self.creg & ~bit) | (m.astype(np.uint64) << np.uint64(creg_bit))
        self.outcomes.append(m.astype(np.int8))
        return m

    def read_outcomes(self):
        """(number of measurements, batch) int8: every outcome in draw order."""
        return np.array(self.outcomes, dtype=np.int8).reshape(len(self.outcomes), self.batch)

    # ------------------------------------------------------------------ read-out
    def flush(self):
        pass

    def read_creg(self, bit):
        return ((self.creg >> np.uint64(bit)) & np.uint64(1)).astype(np.int8)

    def histogram(self, cbits):
        """Counts over shots of the integer formed by the listed creg bits (first = MSB)."""
        idx = np.zeros((self.batch,), dtype=np.int64)
        for b in cbits:
            idx = (idx << 1) | ((self.creg >> np.uint64(b)) & np.uint64(1)).astype(np.int64)
        return np.bincount(idx, minlength=2 ** len(cbits)).astype(np.int64)

    def reduced_dm(self, axes, shot=0):
        """Reduced density matrix of the listed axes (list order = big-endian), one shot."""
        axes = list(axes)
        k = self.n_live
        st = self.state[shot]
        if self.formalism == KET:
            keep = [self.live.index(a) for a in axes]
            rest = [i for i in range(k) if i not in keep]
            psi = np.transpose(st, keep + rest).reshape(2 ** len(keep), -1)
            return psi @ psi.conj().T
        keep = [self.live.index(a) for a in axes]
        rest = [i for i in range(k) if i not in keep]
        rows = list(range(k))
        cols = [k + i if i in keep else i for i in range(k)]  # trace: rest share labels
        out = [i for i in keep] + [k + i for i in keep]
        r = np.einsum(st, rows + cols, out)
        return r.reshape(2 ** len(keep), 2 ** len(keep))

    def fidelity_ket(self, axes, ket, shot=0):
        """<phi| rho_S |phi>  (``qapi.fidelity(qubits, ket, squared=True)``)."""
        v = np.asarray(ket, dtype=complex).reshape(-1)
        return float(np.real(v.conj() @ self.reduced_dm(axes, shot) @ v))

    def full_state(self, shot=None):
        """Embed the live register in the engine's fixed 2^max_qubits layout (free axes |0>),
        with engine axis ``a`` = bit ``a`` of the flat index -- for direct comparison with the
        CUDA engine's device buffer."""
        n = self.max_qubits
        sl = slice(None) if shot is None else slice(shot, shot + 1)
        st = self.state[sl]
        b = st.shape[0]
        k = self.n_live
        if self.formalism == KET:
            out = np.zeros((b,) + (2,) * n, dtype=complex)
            # tensor dim for bit a in a C-order (2,)*n array is n-1-a
            idx = [slice(None)] + [0] * n
            perm_dims = []
            for a in self.live:
                idx[1 + (n - 1 - a)] = slice(None)
            view = out[tuple(idx)]
            # view dims are ordered by decreasing axis id
            order = sorted(self.live, reverse=True)
            perm = [0] + [1 + self.live.index(a) for a in order]
            view[...] = np.transpose(st, perm)
            del perm_dims
            return out.reshape(b, 2 ** n)
        out = np.zeros((b,) + (2,) * (2 * n), dtype=complex)
        idx = [slice(None)] + [0] * (2 * n)
        for a in self.live:
            idx[1 + (n - 1 - a)] = slice(None)
            idx[1 + n + (n - 1 - a)] = slice(None)
        view = out[tuple(idx)]
        order = sorted(self.live, reverse=True)
        perm = [0] + [1 + self.live.index(a) for a in order] + [1 + k + self.live.index(a) for a in order]
        view[...] = np.transpose(st, perm)
        return out.reshape(b, 2 ** n, 2 ** n)
