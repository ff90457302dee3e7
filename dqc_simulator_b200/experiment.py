"""Shot-batch driver: the call a user makes to run a compiled, partitioned circuit for a batch of
noisy shots on one GPU and get per-shot results back on the host.

Mirrors ``examples/run_experiment.py:10-76`` of the reference (``take_experimental_shot`` +
fidelity against the ideal run), with the one-subprocess-per-shot loop the reference documents
(``docs/source/getting_started.rst:683-770``) replaced by one batched launch sequence.
"""
import os

import numpy as np

from . import executor, hardware, opstream

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# examples/run_experiment.py:80-87 + examples/setup_hardware.py:18-33
C2_NOISE = dict(p_depolar_error_cnot=1e-3, single_qubit_gate_error_prob=2e-5, meas_error_prob=3e-3,
                comm_qubit_depolar_rate=0.055, proc_qubit_depolar_rate=0.055)


def c2_hardware(noisy=True, F_werner=0.99, **extra):
    """3 all-to-all NoisyQPUs, 10 positions, 2 comm qubits, ebits at 182 Hz
    (``examples/setup_hardware.py:11-55``)."""
    kw = dict(C2_NOISE) if noisy else {}
    kw.update(extra)
    state = hardware.werner_state(F_werner if noisy else 1.0)
    return hardware.DQCSpec.uniform(3, state4distribution=state, num_positions=10, num_comm_qubits=2, **kw)


def load_c2(circuit):
    return opstream.load(os.path.join(GOLDEN, f"c2_{circuit}5_3qpu_cat.json"))


def ideal_output_ket(qpu_ops, make_backend, dqc=None, max_qubits=7):
    """Ideal-hardware run of the circuit (``take_experimental_shot(circuit)`` with all noise 0):
    returns (data qubits [(pos, node)], their pure output state as a ket, big-endian)."""
    dqc = dqc or c2_hardware(noisy=False)
    be = make_backend("ket", max_qubits, 1)
    ex = executor.DQCExecutor(qpu_ops, dqc, be, forced_outcomes=[0] * 4096).run()
    data = ex.processing_qubits()
    rho = be.reduced_dm(ex.axes_of(data), 0)
    lam, vec = np.linalg.eigh(rho)
    if abs(lam[-1] - 1.0) > 1e-9:
        raise RuntimeError("ideal output is not pure")
    return data, np.ascontiguousarray(vec[:, -1])


class ShotBatchResult:
    __slots__ = ("fidelity", "creg", "executor")

    def __init__(self, fidelity, creg, ex):
        self.fidelity = fidelity
        self.creg = creg
        self.executor = ex


def run_shot_batch(qpu_ops, dqc, engine, data_qubits, ideal_ket, before_flush=None, after_launch=None):
    """One batch of shots: reset the register, walk the compiled op lists (enqueue only), flush the
    gate queue as fused passes, then read per-shot fidelity <phi|rho_S|phi> and the classical
    registers back to the host."""
    engine.reinit()
    ex = executor.DQCExecutor(qpu_ops, dqc, engine).run()
    axes = ex.axes_of(data_qubits)
    if before_flush:
        before_flush()
    engine.flush()
    if after_launch:
        after_launch()
    fid = engine.fidelity_ket_all(axes, ideal_ket)
    creg = engine.read_creg_words()
    return ShotBatchResult(fid, creg, ex)
