"""Independent textbook simulation of an OpenQASM 2 / qelib1.inc circuit (state vector, NumPy).

Used only by ``tests/golden/make_golden.py`` (build container): it simulates the reference's MQT-bench test
circuits straight from their QASM text with the standard qelib1 gate definitions, and the resulting state is
stored next to the op stream the reference's front end compiled from the same file.  The tests then check
that front end + executor + state engine reproduce it (up to a global phase) -- a pin that does not pass
through any of the reference's gate macros.  Qubit q[0] is the most significant factor of the returned ket
(the order of ``qapi.reduced_dm([q0, q1, ...])``)."""
import re

import numpy as np


def u3(t, p, l):
    c, s = np.cos(t / 2), np.sin(t / 2)
    return np.array([[c, -np.exp(1j * l) * s], [np.exp(1j * p) * s, np.exp(1j * (p + l)) * c]])


H = np.array([[1, 1], [1, -1]]) / np.sqrt(2)
X = np.array([[0, 1], [1, 0]], dtype=complex)
ONE_Q = {
    "h": lambda: H, "x": lambda: X, "z": lambda: np.diag([1, -1]).astype(complex),
    "y": lambda: np.array([[0, -1j], [1j, 0]]),
    "s": lambda: np.diag([1, 1j]), "sdg": lambda: np.diag([1, -1j]),
    "t": lambda: np.diag([1, np.exp(1j * np.pi / 4)]), "tdg": lambda: np.diag([1, np.exp(-1j * np.pi / 4)]),
    "u3": u3, "u": u3, "u2": lambda p, l: u3(np.pi / 2, p, l), "u1": lambda l: np.diag([1, np.exp(1j * l)]),
    "p": lambda l: np.diag([1, np.exp(1j * l)]),
    "rx": lambda t: np.array([[np.cos(t / 2), -1j * np.sin(t / 2)], [-1j * np.sin(t / 2), np.cos(t / 2)]]),
    "ry": lambda t: np.array([[np.cos(t / 2), -np.sin(t / 2)], [np.sin(t / 2), np.cos(t / 2)]]),
    "rz": lambda t: np.diag([np.exp(-0.5j * t), np.exp(0.5j * t)]),
}


class Sim:
    """reference_macros=True replays the two places where the reference's QASM macros differ from qelib1.inc
    (``qlib/macros4parsing.py:61`` sets ``minus = ""``): cu1 / cp put all three
    phases, each +lambda/2, on the TARGET (``:386-463``); crx uses u3(+lambda/2, 0, 0) between the CNOTs
    (``:272-307``).  That is the behaviour a drop-in has to reproduce, not fix."""

    def __init__(self, n, reference_macros=False):
        self.reference_macros = reference_macros
        self.n = n
        self.psi = np.zeros((2,) * n, dtype=complex)
        self.psi[(0,) * n] = 1.0

    def g1(self, m, q):
        self.psi = np.moveaxis(np.tensordot(np.asarray(m, dtype=complex), self.psi, axes=([1], [q])), 0, q)

    def ctrl(self, m, c, t):   # |1><1|_c (x) m_t
        idx = [slice(None)] * self.n
        idx[c] = 1
        sub = self.psi[tuple(idx)]
        tt = t - 1 if t > c else t
        self.psi[tuple(idx)] = np.moveaxis(np.tensordot(np.asarray(m, dtype=complex), sub, axes=([1], [tt])), 0, tt)

    def cx(self, c, t):
        self.ctrl(X, c, t)

    def apply(self, name, params, q):
        if name in ONE_Q:
            self.g1(ONE_Q[name](*params), q[0])
        elif name == "cx":
            self.cx(*q)
        elif name == "cz":
            self.ctrl(np.diag([1, -1]), *q)
        elif name in ("cp", "cu1") and self.reference_macros:
            a, b = q
            for _ in range(2):
                self.g1(ONE_Q["u1"](params[0] / 2), b); self.cx(a, b)
            self.g1(ONE_Q["u1"](params[0] / 2), b)
        elif name in ("cp", "cu1"):
            self.ctrl(np.diag([1, np.exp(1j * params[0])]), *q)
        elif name == "crx" and self.reference_macros:
            a, b = q
            self.g1(ONE_Q["u1"](np.pi / 2), b); self.cx(a, b); self.g1(u3(params[0] / 2, 0, 0), b); self.cx(a, b)
            self.g1(u3(params[0] / 2, -np.pi / 2, 0), b)
        elif name == "ch":
            self.ctrl(H, *q)
        elif name == "crx":
            self.ctrl(ONE_Q["rx"](*params), *q)
        elif name == "cu3":
            self.ctrl(u3(*params), *q)
        elif name == "swap":
            self.cx(q[0], q[1]); self.cx(q[1], q[0]); self.cx(q[0], q[1])
        elif name == "rzz":    # qelib1.inc: cx a,b; u1(theta) b; cx a,b
            self.cx(q[0], q[1]); self.g1(ONE_Q["u1"](params[0]), q[1]); self.cx(q[0], q[1])
        elif name == "ccx":    # qelib1.inc decomposition
            a, b, c = q
            for gate, par, qs in (("h", [], [c]), ("cx", [], [b, c]), ("tdg", [], [c]), ("cx", [], [a, c]), ("t", [], [c]),
                                  ("cx", [], [b, c]), ("tdg", [], [c]), ("cx", [], [a, c]), ("t", [], [b]), ("t", [], [c]),
                                  ("h", [], [c]), ("cx", [], [a, b]), ("t", [], [a]), ("tdg", [], [b]), ("cx", [], [a, b])):
                self.apply(gate, par, qs)
        elif name == "rccx":   # qelib1.inc: relative-phase Toffoli
            a, b, c = q
            pi = np.pi
            for gate, par, qs in (("u2", [0, pi], [c]), ("u1", [pi / 4], [c]), ("cx", [], [b, c]), ("u1", [-pi / 4], [c]),
                                  ("cx", [], [a, c]), ("u1", [pi / 4], [c]), ("cx", [], [b, c]), ("u1", [-pi / 4], [c]),
                                  ("u2", [0, pi], [c])):
                self.apply(gate, par, qs)
        else:
            raise NotImplementedError(name)


def simulate(text, reference_macros=False):
    """-> ket of all qubits, registers concatenated in declaration order, first qubit most significant."""
    regs, n = {}, 0
    for name, size in re.findall(r"qreg\s+(\w+)\[(\d+)\]", text):
        regs[name] = n
        n += int(size)
    sim = Sim(n, reference_macros)
    for raw in text.splitlines():
        line = raw.split("//")[0].strip().rstrip(";")
        if not line or line.split()[0] in ("OPENQASM", "include", "qreg", "creg", "barrier", "measure"):
            continue
        m = re.match(r"(\w+)\s*(?:\(([^)]*)\))?\s+(.*)$", line)
        name, par, args = m.group(1), m.group(2), m.group(3)
        params = [float(eval(p, {"__builtins__": {}}, {"pi": np.pi})) for p in par.split(",")] if par else []
        qs = [regs[r] + int(i) for r, i in re.findall(r"(\w+)\[(\d+)\]", args)]
        sim.apply(name, params, qs)
    return sim.psi.reshape(-1)
