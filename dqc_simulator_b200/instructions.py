"""Instruction tokens and operator matrices at the QuantumProgram boundary.

Host-side mirror of the names the reference passes to ``QuantumProgram.apply``
(``software/dqc_control.py:191-196``): NetSquid's ``netsquid.components.instructions``
built-ins and the extra gates of ``qlib/gates.py:96-220``.  A gate
tuple's first element is either a bare instruction or an ``(instruction, operator)`` pair
(``dqc_control.py:793-801``); :func:`resolve` turns either form -- ours, or a foreign
duck-typed object carrying ``.name`` and ``.operator.arr`` such as a real NetSquid
``IGate`` -- into ``(kind, num_qubits, matrix, physical_class)``.

Matrix conventions (Appendix A): big-endian qubit order, ``ctrl`` = control on the first
index, ``INSTR_U`` in the SU(2) form of ``qlib/gates.py:145-149``.
"""
import numpy as np

_S2 = 1.0 / np.sqrt(2.0)

I2 = np.eye(2, dtype=complex)
X = np.array([[0, 1], [1, 0]], dtype=complex)
Y = np.array([[0, -1j], [1j, 0]], dtype=complex)
Z = np.array([[1, 0], [0, -1]], dtype=complex)
H = np.array([[_S2, _S2], [_S2, -_S2]], dtype=complex)
S = np.array([[1, 0], [0, 1j]], dtype=complex)
T = np.array([[1, 0], [0, np.exp(0.25j * np.pi)]], dtype=complex)
SWAP = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=complex)


def ctrl(u):
    """|0><0| (x) I + |1><1| (x) U, control first (``Operator.ctrl``)."""
    u = np.asarray(u, dtype=complex)
    d = u.shape[0]
    out = np.eye(2 * d, dtype=complex)
    out[d:, d:] = u
    return out


def u_matrix(theta, phi, lam):
    """``INSTR_U(theta, phi, lambda)`` operator, ``qlib/gates.py:145-149``."""
    c, s = np.cos(theta / 2), np.sin(theta / 2)
    return np.array(
        [[np.exp(-0.5j * (phi + lam)) * c, -np.exp(-0.5j * (phi - lam)) * s],
         [np.exp(0.5j * (phi - lam)) * s, np.exp(0.5j * (phi + lam)) * c]], dtype=complex)


def rz_matrix(angle):
    """``create_rotation_op(angle, (0,0,1))`` = cos(a/2) I - i sin(a/2) Z (``gates.py:208``)."""
    return np.array([[np.exp(-0.5j * angle), 0], [0, np.exp(0.5j * angle)]], dtype=complex)


def state_gen_matrix(alpha, beta):
    """Non-unitary ``make_state_gen_op`` (``qlib/gates.py:45-48``)."""
    return np.diag([alpha, beta]).astype(complex) @ np.array([[1, 1], [1, -1]], dtype=complex)


class Instruction:
    """Token with the attributes the executor needs: a ``name``, a ``kind`` in
    {"init", "measure", "discard", "gate"}, ``num_qubits`` and an optional fixed ``matrix``.
    ``phys`` names the row of the physical-instruction table that supplies duration and
    noise (``hardware/quantum_processors.py:344-456``)."""

    def __init__(self, name, kind="gate", num_qubits=1, matrix=None, phys=None):
        self.name = name
        self.kind = kind
        self.num_qubits = num_qubits
        self.matrix = None if matrix is None else np.asarray(matrix, dtype=complex)
        self.phys = phys or name

    def __repr__(self):
        return f"Instruction({self.name})"


INSTR_INIT = Instruction("init", kind="init")
INSTR_MEASURE = Instruction("measure", kind="measure")
INSTR_DISCARD = Instruction("discard", kind="discard")
INSTR_H = Instruction("h_gate", matrix=H)
INSTR_X = Instruction("x_gate", matrix=X)
INSTR_Y = Instruction("y_gate", matrix=Y)
INSTR_Z = Instruction("z_gate", matrix=Z)
INSTR_S = Instruction("s_gate", matrix=S)
INSTR_T = Instruction("t_gate", matrix=T)
INSTR_IDENTITY = Instruction("I", matrix=I2)
INSTR_S_DAGGER = Instruction("S_dagger", matrix=S.conj().T)
INSTR_T_DAGGER = Instruction("T_dagger", matrix=T.conj().T)
INSTR_CNOT = Instruction("cnot_gate", num_qubits=2, matrix=ctrl(X))
INSTR_CZ = Instruction("cz_gate", num_qubits=2, matrix=ctrl(Z))
INSTR_CS = Instruction("cs_gate", num_qubits=2, matrix=ctrl(S))
INSTR_CH = Instruction("CH", num_qubits=2, matrix=ctrl(H))
INSTR_CT = Instruction("CT", num_qubits=2, matrix=ctrl(T))
INSTR_CY = Instruction("CY", num_qubits=2, matrix=ctrl(Y))
INSTR_SWAP = Instruction("swap_gate", num_qubits=2, matrix=SWAP)
# parameter-less placeholders whose operator travels with the gate tuple (gates.py:110-115)
INSTR_SINGLE_QUBIT_UNITARY = Instruction("single_qubit_unitary")
INSTR_TWO_QUBIT_UNITARY = Instruction("two_qubit_unitary", num_qubits=2)
INSTR_SINGLE_QUBIT_NEGLIGIBLE_TIME = Instruction("neglibible_time_instr")
INSTR_TWO_QUBIT_NEGLIGIBLE_TIME = Instruction("neglible_time_2_qubit", num_qubits=2)

_ALL = [v for v in list(globals().values()) if isinstance(v, Instruction)]
BY_NAME = {i.name: i for i in _ALL}
# spellings used by NetSquid's own instruction objects [NS-doc] and by our fixture stub
BY_NAME.update({
    "init_qubit": INSTR_INIT, "measure_qubit": INSTR_MEASURE, "discard_qubit": INSTR_DISCARD,
    "cx_gate": INSTR_CNOT, "i_gate": INSTR_IDENTITY,
})


def INSTR_U(theta, phi, lam, controlled=False):
    """(instruction, operator) pair like ``qlib/gates.py:118-157``."""
    m = u_matrix(theta, phi, lam)
    if controlled:
        return (INSTR_TWO_QUBIT_UNITARY, ctrl(m))
    return (INSTR_SINGLE_QUBIT_UNITARY, m)


def instrNop_RZ(angle, controlled=False, conjugate=False):
    """``qlib/gates.py:181-210``."""
    m = rz_matrix(angle)
    if controlled:
        m = ctrl(m)
    if conjugate:
        m = m.conj().T
    return (INSTR_TWO_QUBIT_UNITARY if controlled else INSTR_SINGLE_QUBIT_UNITARY, m)


instrNop_SX = (INSTR_SINGLE_QUBIT_UNITARY, S.conj().T @ H @ S.conj().T)
instrNop_SXDG = (INSTR_SINGLE_QUBIT_UNITARY, S @ H @ S)
instrNop_CSX = (INSTR_TWO_QUBIT_UNITARY, ctrl(instrNop_SX[1]))


def _matrix_of(op):
    if op is None:
        return None
    if hasattr(op, "arr"):  # NetSquid-style Operator
        op = op.arr
    return np.asarray(op, dtype=complex)


def resolve(gate):
    """gate-tuple head -> (token, matrix).  Accepts a bare instruction or an
    (instruction, operator) pair (``dqc_control.py:793-801``)."""
    op = None
    if isinstance(gate, (tuple, list)):
        gate, op = gate[0], gate[1]
    if isinstance(gate, Instruction):
        tok = gate
    else:
        name = getattr(gate, "name", None)
        if name not in BY_NAME:
            raise KeyError(f"unknown instruction {gate!r}")
        tok = BY_NAME[name]
        if op is None and tok.matrix is None:
            op = getattr(gate, "operator", None) or getattr(gate, "_operator", None)
    m = _matrix_of(op)
    if m is None:
        m = tok.matrix
    if tok.kind == "gate" and m is None:
        raise ValueError(f"{tok.name} needs an operator (gates.py:110-115)")
    return tok, m
