"""Level-1 boundary: a NetSquid ``QRepr``-shaped quantum state whose arithmetic runs in
libdqcengine.

NetSquid picks the state representation through ``ns.set_qstate_formalism(QFormalism.X)``; the enum values are
repr classes, and every ``qubitapi`` call (``operate``, ``measure``, ``stochastic_operate``, ``apply_pauli_noise``,
``depolarize``, ``reduced_dm``, ``fidelity``, ``combine`` ...) lands on a method of the repr object that holds
the qubits' joint state.  The reference reaches into that object directly as well: ``qstate.qrepr``,
``.dm`` / ``.ket``, ``.num_qubits`` (``hardware/noise_models.py:92, 102, 107, 129``, ``util/helper.py:210``).

``B200Repr`` is that object on top of one engine handle (one joint state = one handle, batch 1): the method names
and argument meanings are NetSquid's [NS-doc], the body of each is one or two C-ABI calls.  With NetSquid importable
(Python 3.9) it subclasses ``netsquid.qubits.qrepr.QRepr`` and ``ns.set_qstate_formalism(B200DMRepr)`` routes the
unmodified ``dqc_simulator`` to the GPU; here it is exercised through a NetSquid-free harness
(``tests/test_compat_qrepr.py``) that replays the call sequences the reference makes.
Batching over shots is the level-2 executor's job (NetSquid's DES advances one trajectory at a time)."""
import numpy as np

try:  # pragma: no cover -- only with a real NetSquid (closed cp39 wheel)
    from netsquid.qubits.qrepr import QRepr as _Base
except Exception:  # noqa: BLE001
    class _Base:
        """Placeholder base where NetSquid is absent."""


_PAULI = [np.eye(2, dtype=complex), np.array([[0, 1], [1, 0]], dtype=complex),
          np.array([[0, -1j], [1j, 0]], dtype=complex), np.diag([1.0 + 0j, -1.0])]


def _matrix(operator):
    """NetSquid ``Operator`` (``.arr``), ``(instruction, operator)`` pairs of qlib/gates.py, or a plain array."""
    arr = getattr(operator, "arr", None)
    if arr is None and hasattr(operator, "operator"):
        arr = operator.operator.arr
    return np.asarray(operator if arr is None else arr, dtype=complex)


class B200Repr(_Base):
    """Joint state of ``num_qubits`` qubits (index i of NetSquid's ``indices`` = i-th qubit of the state, big-endian
    in ``ket`` / ``dm`` like ``kron(q0, q1, ...)``) held by one engine handle."""

    formalism = "dm"

    def __init__(self, num_qubits=1, max_qubits=None, engine_factory=None, seed=0, **_ignored):
        if engine_factory is None:
            from ..engine import Engine as engine_factory  # noqa: N813
        self.max_qubits = int(max_qubits or max(num_qubits, 8))
        self._eng = engine_factory(self.formalism, self.max_qubits, 1, seed=seed)
        self._axes = []
        for _ in range(num_qubits):
            self._new_axis()

    # ---- bookkeeping the reference reads directly --------------------------------------------------------------
    @property
    def qrepr(self):            # qstate.qrepr (noise_models.py:129): the repr is its own state object here
        return self

    @property
    def num_qubits(self):
        return len(self._axes)

    @property
    def indices(self):          # positions of the qubits inside the joint state (noise_models.py:92-107)
        return list(range(len(self._axes)))

    def _new_axis(self):
        free = [a for a in range(self.max_qubits) if a not in self._axes]
        if not free:
            raise ValueError("B200Repr: register full (raise max_qubits)")
        ax = free[0]
        self._eng.alloc_axis(ax)
        self._axes.append(ax)
        return ax

    def _ax(self, indices):
        return [self._axes[int(i)] for i in indices]

    # ---- qubitapi.operate(qubits, operator) : QuantumProgram gates, noise_models.py:364 -----------------------------
    def operate(self, operator, indices):
        m = _matrix(operator)
        ax = self._ax(indices)
        if len(ax) == 1:
            self._eng.apply_1q(ax[0], m)
        elif len(ax) == 2:
            self._eng.apply_2q(ax[0], ax[1], m)
        else:
            self._eng.apply_kq(ax, m)

    # ---- qubitapi.stochastic_operate(qubits, operators, p_weights) -------------------------------------------------
    def stochastic_operate(self, operators, p_weights, indices):
        """DM formalism: the mixture sum_i p_i O_i rho O_i^dagger, applied analytically (Appendix A).  One- and
        two-qubit Pauli mixtures (what the reference builds, noise_models.py:343-366) map to closed-form channels;
        anything else runs as a one-qubit Kraus channel."""
        ax = self._ax(indices)
        mats = [_matrix(o) for o in operators]
        w = np.asarray(p_weights, dtype=float)
        if len(ax) == 1:
            probs = np.zeros(4)
            for m, p in zip(mats, w):
                k = next((i for i, P in enumerate(_PAULI) if np.allclose(m, P)), None)
                if k is None:
                    return self._eng.kraus_1q(ax[0], [np.sqrt(p) * m for m, p in zip(mats, w)])
                probs[k] += p
            return self._eng.pauli_channel(ax[0], probs)
        if len(ax) == 2 and len(mats) == 16 and np.allclose(w[1:], w[1]):
            return self._eng.depolarize(ax, 16.0 * w[1])      # (1 - 15p/16, p/16 ...) == NDimDepolarNoiseModel
        if len(ax) == 2:
            # any mixture of two-qubit Pauli products (correlated noise): one k = 2 Pauli channel
            # (dqe_enqueue_pauli_channel_k: a single 16 x 16 superoperator block in the dm formalism)
            products = [np.kron(a, b) for a in _PAULI for b in _PAULI]
            probs = np.zeros(16)
            for m, p in zip(mats, w):
                k = next((i for i, P in enumerate(products) if m.shape == (4, 4) and np.allclose(m, P)), None)
                if k is None:
                    raise NotImplementedError("stochastic_operate: two-qubit operators must be Pauli products")
                probs[k] += p
            return self._eng.pauli_channel(ax, probs)
        raise NotImplementedError("stochastic_operate: unsupported operator set")

    def apply_pauli_noise(self, index, probs):
        """(pI, pX, pY, pZ) on one qubit: MeasurementNoiseModel (noise_models.py:295-307), dephasing."""
        self._eng.pauli_channel(self._axes[int(index)], np.asarray(probs, dtype=float))

    def depolarize(self, indices, prob):
        """NetSquid DepolarNoiseModel (one qubit) / k-qubit depolarising (noise_models.py:33-131, 343-366)."""
        self._eng.depolarize(self._ax(indices), float(prob))

    def dephase(self, index, prob):
        self._eng.pauli_channel(self._axes[int(index)], [1.0 - prob, 0.0, 0.0, prob])

    def amplitude_dampen(self, index, gamma, prob=1.0):
        g = float(gamma)
        self._eng.kraus_1q(self._axes[int(index)], [np.array([[1, 0], [0, np.sqrt(1 - g)]], dtype=complex),
                                                    np.array([[0, np.sqrt(g)], [0, 0]], dtype=complex)])

    # ---- qubitapi.measure(qubit, observable=Z, discard=...) : INSTR_MEASURE ------------------------------------------
    def measure(self, index, observable=None, meas_operators=None, discard=False, rng=None, forced=None):
        """Z-basis measurement; returns (outcome, probability of that outcome) like NetSquid [NS-doc]."""
        if observable is not None and not np.allclose(_matrix(observable), _PAULI[3]):
            raise NotImplementedError("only Z-basis measurements are on the path (dqc_control.py:298, 366, 446-459)")
        ax = self._axes[int(index)]
        p1 = float(np.real(self._eng.reduced_dm([ax], 0)[1, 1]))
        self._eng.measure(ax, 0, -1 if forced is None else int(forced), bool(discard), 0.0)
        m = int(self._eng.read_creg(0)[0])
        if discard:
            self._axes.pop(int(index))
        return m, (p1 if m else 1.0 - p1)

    # ---- state growth / shrink: QState merge on the first multi-qubit op, discard ------------------------------------
    def add_qubits(self, n=1):
        return [self.indices[-1] for _ in range(n) if self._new_axis() is not None]

    def combine(self, other):
        """Merge another (small) repr into this one: tensor product, other's qubits appended.  The engine keeps
        one joint register from the start, so this is only needed when NetSquid created a separate QState
        (the 2-qubit ebit of a QSource, connections.py:145-150): its state is injected onto two fresh axes."""
        k = other.num_qubits
        if k != 2:
            raise NotImplementedError("combine: single fresh qubits are add_qubits(); pairs are injected")
        a, b = self._new_free(), None
        b = self._new_free(exclude=[a])
        if self.formalism == "dm":
            self._eng.inject_2q(a, b, other.dm, True)
        else:
            self._eng.inject_2q(a, b, other.ket, False)
        self._axes += [a, b]

    def _new_free(self, exclude=()):
        return next(a for a in range(self.max_qubits) if a not in self._axes and a not in exclude)

    def inject_pair(self, state):
        """Fresh entangled pair (``state4distribution``: ket (4,) or dm (4, 4)) on two new qubits
        (physical_layer.py:451-479)."""
        state = np.asarray(state, dtype=complex)
        a = self._new_free()
        b = self._new_free(exclude=[a])
        self._eng.inject_2q(a, b, state, state.size == 16)
        self._axes += [a, b]
        return [len(self._axes) - 2, len(self._axes) - 1]

    def drop_qubit(self, index):
        """Discard a qubit (partial trace in the dm formalism)."""
        self._eng.free_axis(self._axes.pop(int(index)))

    # ---- read-out: helper.py:151, 209-214; run_experiment.py:74-75 ---------------------------------------------------
    def reduced_dm(self, indices=None):
        idx = self.indices if indices is None else list(indices)
        return self._eng.reduced_dm(self._ax(idx), 0)

    @property
    def dm(self):
        return self.reduced_dm()

    @property
    def ket(self):
        if self.formalism != "ket":
            raise ValueError("a density-matrix repr has no ket")
        n = len(self._axes)
        psi = self._eng.full_state(0)[0] if hasattr(self._eng, "full_state") else None
        full = np.asarray(psi).reshape((2,) * self.max_qubits)           # axis a = bit a: dim index = max-1-a
        perm = [self.max_qubits - 1 - a for a in self._axes]
        rest = [d for d in range(self.max_qubits) if d not in perm]
        return np.transpose(full, perm + rest).reshape(2 ** n, -1)[:, 0].copy()

    def fidelity(self, indices, reference, squared=True):
        """``qapi.fidelity(qubits, reference, squared=True)``: <phi|rho|phi> for a ket reference."""
        from ..experiment import fidelity_squared

        f2 = fidelity_squared(self.reduced_dm(indices), np.asarray(reference, dtype=complex))
        return f2 if squared else float(np.sqrt(f2))


class B200DMRepr(B200Repr):
    formalism = "dm"


class B200KetRepr(B200Repr):
    formalism = "ket"
