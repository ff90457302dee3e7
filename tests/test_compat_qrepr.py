"""Level-1 boundary: ``compat.qrepr.B200Repr`` driven the way NetSquid's ``qubitapi`` and the
reference's noise models drive a ``QRepr`` -- ``operate / stochastic_operate / apply_pauli_noise / depolarize /
measure / combine / reduced_dm / fidelity`` and the attributes the reference reads directly (``.qrepr``, ``.dm``,
``.num_qubits``, ``.indices``: hardware/noise_models.py:92-130, util/helper.py:209-214).  Expected values are plain
NumPy; the backend is the oracle on CPU and the CUDA engine (through the C-ABI) on the GPU."""
import functools as ft

import numpy as np
import pytest

from dqc_simulator_b200.compat.qrepr import B200DMRepr, B200KetRepr

I2 = np.eye(2, dtype=complex)
X = np.array([[0, 1], [1, 0]], dtype=complex)
Y = np.array([[0, -1j], [1j, 0]], dtype=complex)
Z = np.diag([1.0 + 0j, -1.0])
H = np.array([[1, 1], [1, -1]], dtype=complex) / np.sqrt(2)
CNOT = np.eye(4, dtype=complex)[[0, 1, 3, 2]]
B00 = np.array([1, 0, 0, 1], dtype=complex) / np.sqrt(2)


class Op:            # the only thing the adapter needs of a NetSquid Operator
    def __init__(self, arr):
        self.arr = np.asarray(arr, dtype=complex)


def _oracle_factory(*a, **k):
    from oracle.numpy_engine import OracleEngine

    return OracleEngine(*a, **k)


def _cuda_factory(*a, **k):
    from dqc_simulator_b200.engine import Engine

    return Engine(*a, **k)


BACKENDS = [pytest.param(_oracle_factory, id="oracle"), pytest.param(_cuda_factory, id="cuda", marks=pytest.mark.gpu)]


def werner(F):
    s = 2 ** -0.5
    b = [np.array(v, dtype=complex) for v in ([s, 0, 0, s], [0, s, s, 0], [s, 0, 0, -s], [0, s, -s, 0])]
    return F * np.outer(b[0], b[0].conj()) + sum((1 - F) / 3 * np.outer(v, v.conj()) for v in b[1:])


@pytest.mark.parametrize("factory", BACKENDS)
def test_noise_model_call_sequences(factory):
    """NDimDepolarNoiseModel.error_operation (noise_models.py:343-366: 16 Pauli products, weights
    1 - 15p/16, p/16 ...), apply_analytical_depolarisation (:33-131: reads .dm, depolarises) and
    MeasurementNoiseModel (:295-307) on a Bell pair with a spectator."""
    p, e = 0.3, 0.1
    r = B200DMRepr(num_qubits=3, max_qubits=5, engine_factory=factory)
    assert r.qrepr is r and r.num_qubits == 3 and r.indices == [0, 1, 2]
    r.operate(Op(H), [0])
    r.operate(Op(CNOT), [0, 1])
    r.operate(Op(H), [2])
    paulis = [Op(ft.reduce(np.kron, (a, b))) for a in (I2, X, Y, Z) for b in (I2, X, Y, Z)]
    r.stochastic_operate(paulis, [1 - 15 * p / 16] + [p / 16] * 15, [0, 1])
    bell = np.outer(B00, B00.conj())
    plus = np.full((2, 2), 0.5, dtype=complex)
    want = np.kron((1 - p) * bell + p * np.eye(4) / 4, plus)
    assert np.abs(r.dm - want).max() < 1e-12
    r.depolarize([0, 1], p)                                   # AnalyticalDepolarisationModel: same channel again
    want01 = (1 - p) * ((1 - p) * bell + p * np.eye(4) / 4) + p * np.eye(4) / 4
    assert np.abs(r.reduced_dm([0, 1]) - want01).max() < 1e-12
    assert np.abs(r.reduced_dm([1, 0]) - want01).max() < 1e-12     # symmetric state: list order is big-endian
    r.apply_pauli_noise(2, (1 - e, e, 0.0, 0.0))             # X flip on |+>: no change
    assert np.abs(r.reduced_dm([2]) - plus).max() < 1e-12
    r.apply_pauli_noise(2, (1 - e, 0.0, 0.0, e))             # dephasing
    assert np.abs(r.reduced_dm([2]) - np.array([[0.5, 0.5 - e], [0.5 - e, 0.5]])).max() < 1e-12
    assert abs(r.fidelity([0, 1], B00) - np.real(B00.conj() @ want01 @ B00)) < 1e-12


@pytest.mark.parametrize("factory", BACKENDS)
@pytest.mark.parametrize("ma,mb", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_cat_comm_remote_cnot_through_the_repr(factory, ma, mb):
    """The cat-entanglement remote CNOT (dqc_control.py:266-403, Appendix B) as qubitapi calls on ONE repr: the ebit
    arrives as a separate two-qubit state and is combined, measurements are discarding, corrections follow the
    forced outcomes.  Control |+>, target |0>  ->  Bell pair for every (ma, mb)."""
    r = B200DMRepr(num_qubits=2, max_qubits=6, engine_factory=factory)
    r.operate(Op(H), [0])                                     # control d = index 0, target t = index 1
    c, c2 = r.inject_pair(werner(1.0))                        # ebit on (c, c')
    r.operate(Op(CNOT), [0, c])
    m, pm = r.measure(c, observable=Op(Z), discard=True, forced=ma)
    assert m == ma and abs(pm - 0.5) < 1e-12
    c2 -= 1                                                   # indices above the discarded qubit shift down
    if ma:
        r.operate(Op(X), [c2])
    r.operate(Op(CNOT), [c2, 1])
    r.operate(Op(H), [c2])
    m, pm = r.measure(c2, discard=True, forced=mb)
    assert m == mb and abs(pm - 0.5) < 1e-12
    if mb:
        r.operate(Op(Z), [0])
    assert r.num_qubits == 2
    assert abs(r.fidelity([0, 1], B00) - 1.0) < 1e-12
    assert np.abs(r.dm - np.outer(B00, B00.conj())).max() < 1e-12


@pytest.mark.parametrize("factory", BACKENDS)
def test_ket_repr_and_generic_channels(factory):
    r = B200KetRepr(num_qubits=3, max_qubits=4, engine_factory=factory)
    rng = np.random.default_rng(3)
    q, rr = np.linalg.qr(rng.normal(size=(8, 8)) + 1j * rng.normal(size=(8, 8)))
    u8 = q * (np.diag(rr) / np.abs(np.diag(rr)))
    r.operate(Op(u8), [2, 0, 1])                             # a dense 3-qubit operator (OP_KBLOCK on the GPU)
    psi = np.zeros(8, dtype=complex); psi[0] = 1
    want = (u8 @ psi).reshape(2, 2, 2).transpose(1, 2, 0).reshape(-1)      # operator order (2, 0, 1) -> state order
    assert np.abs(r.ket - want).max() < 1e-12
    d = B200DMRepr(num_qubits=1, max_qubits=2, engine_factory=factory)
    d.operate(Op(X), [0])
    d.amplitude_dampen(0, 0.25)
    assert np.abs(d.dm - np.diag([0.25, 0.75])).max() < 1e-12
    d.operate(Op(H), [0])
    d.dephase(0, 0.2)
    hz = H @ np.diag([0.25, 0.75]) @ H
    hz[0, 1] *= 0.6; hz[1, 0] *= 0.6
    assert np.abs(d.dm - hz).max() < 1e-12


@pytest.mark.parametrize("factory", BACKENDS)
def test_correlated_two_qubit_pauli_mixture(factory):
    """qubitapi.stochastic_operate with a NON-uniform set of two-qubit Pauli products (correlated XX / ZZ noise on a
    Bell pair): one k = 2 Pauli channel on the engine; expected value by plain NumPy."""
    r = B200DMRepr(num_qubits=2, max_qubits=3, engine_factory=factory)
    r.operate(Op(H), [0])
    r.operate(Op(CNOT), [0, 1])
    ops = [Op(np.kron(I2, I2)), Op(np.kron(X, X)), Op(np.kron(Z, Z)), Op(np.kron(X, Z))]
    w = [0.7, 0.15, 0.1, 0.05]
    r.stochastic_operate(ops, w, [0, 1])
    rho0 = np.outer(B00, B00.conj())
    want = sum(p * o.arr @ rho0 @ o.arr.conj().T for o, p in zip(ops, w))
    assert np.abs(r.dm - want).max() < 1e-12
    with pytest.raises(NotImplementedError, match="Pauli products"):
        r.stochastic_operate([Op(np.kron(H, I2))], [1.0], [0, 1])
