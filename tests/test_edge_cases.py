"""Edge cases of the boundary: the reference's own error paths (dqc_control.py:283-287, 953-974) on the
executor (CPU), and empty / minimal / maximal inputs on the engine (GPU)."""
import numpy as np
import pytest

from dqc_simulator_b200 import executor, hardware
from dqc_simulator_b200 import instructions as ins
from oracle.numpy_engine import OracleEngine

H = np.array([[1, 1], [1, -1]], dtype=complex) / np.sqrt(2)
X = np.array([[0, 1], [1, 0]], dtype=complex)


def test_scheduling_error_when_no_comm_qubit_is_free():
    """Three remote gates in one slice on QPUs with two comm qubits: the reference raises
    'Scheduling error: No comm-qubits free' (dqc_control.py:283-287); 'correct' reserves one comm
    qubit per remote gate until its disentangle_start."""
    ops = {
        "node_0": [[(ins.INSTR_INIT, [2, 3, 4]), (2, "node_1", "cat", "entangle"), (3, "node_1", "cat", "entangle"),
                    (4, "node_1", "cat", "entangle")]],
        "node_1": [[(ins.INSTR_INIT, [2]), ("node_0", "cat", "correct"), ("node_0", "cat", "correct"),
                    ("node_0", "cat", "correct")]],
    }
    dqc = hardware.DQCSpec.uniform(2, num_positions=6, num_comm_qubits=2)
    with pytest.raises(Exception, match="No comm-qubits free"):
        executor.DQCExecutor(ops, dqc, OracleEngine("dm", 8, 1)).run()


def test_unfinished_circuit_is_detected():
    """A 'correct' with no matching 'entangle' never completes: UnfinishedQuantumCircuitError, like
    DQCMasterProtocol.check_quantum_circuit_finished (dqc_control.py:953-974)."""
    ops = {"node_0": [[(ins.INSTR_INIT, [2])]], "node_1": [[(ins.INSTR_INIT, [2]), ("node_0", "cat", "correct")]]}
    dqc = hardware.DQCSpec.uniform(2, num_positions=4, num_comm_qubits=2)
    with pytest.raises(executor.UnfinishedQuantumCircuitError):
        executor.DQCExecutor(ops, dqc, OracleEngine("dm", 6, 1)).run()


def test_gate_on_empty_memory_position_is_an_error():
    ops = {"node_0": [[(ins.INSTR_X, 3)]]}
    dqc = hardware.DQCSpec.uniform(1, num_positions=5)
    with pytest.raises(RuntimeError, match="holds no qubit"):
        executor.DQCExecutor(ops, dqc, OracleEngine("ket", 4, 1)).run()


def test_empty_circuit_runs():
    ex = executor.DQCExecutor({"node_0": [[]], "node_1": []}, hardware.DQCSpec.uniform(2), OracleEngine("ket", 2, 1)).run()
    assert ex.finished and ex.measurements == 0


def test_invalid_measurement_result_message():
    """_apply_if mirrors the reference's ValueError text for outcomes that are not 0/1
    (dqc_control.py:351-359, 396-403)."""
    ex = executor.DQCExecutor({"node_0": [[]]}, hardware.DQCSpec.uniform(1), OracleEngine("ket", 2, 1))
    with pytest.raises(ValueError, match="Measurement result must be an integer with value 0 or 1"):
        ex._apply_if(executor.QuantumProgram(), ins.INSTR_X, 0, 2)


# ---- engine (GPU) ---------------------------------------------------------------------------------
@pytest.mark.gpu
def test_engine_minimal_and_empty_inputs():
    from dqc_simulator_b200.engine import Engine

    e = Engine("ket", 1, 1)
    e.flush(); e.sync()                                    # empty queue
    assert e.full_state(0)[0, 0] == 1 and e.live == []
    e.alloc_axis(0)
    e.apply_1q(0, H)
    e.measure(0, 63, 1, False, 0.0)                        # last creg bit, forced outcome
    assert int(e.read_creg_words()[0]) == 1 << 63
    assert abs(e.full_state(0)[0, 1] - 1) < 1e-15
    e.free_axis(0)
    assert e.live == [] and abs(e.full_state(0)[0, 0] - 1) < 1e-15
    d = Engine("dm", 1, 3)
    d.alloc_axis(0)
    d.apply_1q(0, X)
    d.depolarize([0], 1.0)                                 # fully depolarised
    assert np.allclose(d.reduced_dm([0], 2), np.eye(2) / 2)
    assert d.histogram([0]).tolist() == [3, 0]


@pytest.mark.gpu
def test_large_batch_of_tiny_registers():
    from dqc_simulator_b200.engine import Engine

    e = Engine("ket", 2, 200000, seed=5)
    for a in range(2):
        e.alloc_axis(a)
    e.apply_1q(0, H)
    e.apply_2q(0, 1, np.eye(4)[[0, 1, 3, 2]])
    e.measure(0, 0, -1, False, 0.0)
    e.measure(1, 1, -1, False, 0.0)
    hist = e.histogram([0, 1])
    assert hist[1] == 0 and hist[2] == 0 and hist.sum() == 200000      # Bell pair: outcomes agree
    assert abs(hist[0] - 100000) < 2000


@pytest.mark.gpu
def test_register_capacity_and_argument_errors():
    from dqc_simulator_b200.engine import Engine, EngineError

    e = Engine("dm", 3, 1)
    for a in range(3):
        e.alloc_axis(a)
    with pytest.raises(EngineError, match="out of range"):
        e.alloc_axis(3)
    with pytest.raises(EngineError, match="axes must differ"):
        e.apply_2q(1, 1, np.eye(4))
    with pytest.raises(EngineError, match="occupied"):
        e.inject_2q(0, 1, [1, 0, 0, 0], False)
    with pytest.raises(EngineError, match="k = 1 or 2"):
        e.depolarize([0, 1, 2], 0.1)
    with pytest.raises(EngineError):
        Engine("ket", 41, 1)


def test_handle_limits_and_option_errors_without_a_gpu():
    """C-ABI argument checks that need no device (plan-only handles): a density matrix is at most 16 qubits
    (4^16 complex128 = 64 GiB, and the tile-base arithmetic of the diagonal enumeration is sized for it),
    options reject out-of-range values and unknown keys with a message, never an exception across the ABI."""
    from dqc_simulator_b200 import engine

    with pytest.raises(engine.EngineError, match="dm <= 16"):
        engine.Engine("dm", 17, 1, device=-1)
    engine.Engine("dm", 16, 1, device=-1).close()
    e = engine.Engine("dm", 4, 2, device=-1)
    for key, bad in (("prefetch", -2), ("prefetch", 1 << 21), ("threads", 96)):
        with pytest.raises(engine.EngineError):
            e.set_option(key, bad)
    with pytest.raises(engine.EngineError, match="unknown option"):
        e.set_option("variant", 1)     # the register-budget variants are gone
    for key, ok in (("prefetch", -1), ("prefetch", 0), ("tmap", 0), ("cxperm", 0), ("factor", 0), ("cluster", 0)):
        e.set_option(key, ok)
    e.close()


def test_negligible_time_instructions_have_no_noise_and_three_ns():
    """the reference's two negligible-time tokens (qlib/gates.py:112-115; physical instructions at
    quantum_processors.py:450-455) take 3 ns and carry no noise model, whatever the gate error settings."""
    from dqc_simulator_b200 import hardware, instructions as ins

    spec = hardware.QPUSpec(single_qubit_gate_error_prob=0.01, p_depolar_error_cnot=0.02)
    for tok in (ins.INSTR_SINGLE_QUBIT_NEGLIGIBLE_TIME, ins.INSTR_TWO_QUBIT_NEGLIGIBLE_TIME):
        ph = spec.phys(tok)
        assert ph.duration == 3.0 and ph.noise == "none" and ph.prob == 0.0
    assert spec.phys(ins.INSTR_H).duration == 135e3 and spec.phys(ins.INSTR_H).noise == "depol1"


def test_planner_limits_at_the_edge(capfd, monkeypatch):
    """The silent limits of the planner, on a plan-only handle: a run of ops longer than the 48-micro-op budget of a
    pass is CUT into passes that each respect it (never an error, never an over-full pass); the classical register
    has exactly 64 bits (bit 63 works, bit 64 is refused with a message)."""
    from dqc_simulator_b200 import engine

    rng = np.random.default_rng(0)

    def ru(d):
        q, _ = np.linalg.qr(rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d)))
        return q

    cnot = np.eye(4)[[0, 1, 3, 2]]
    monkeypatch.setenv("DQE_DUMP_SMEM", "1")
    for formalism, n in (("ket", 9), ("dm", 5)):
        e = engine.Engine(formalism, n, 2, device=-1)
        for a in range(n):
            e.alloc_axis(a)
        for i in range(400):                       # nothing here folds away: a dense gate, then a CNOT across axes
            e.apply_1q(i % n, ru(2))
            e.apply_2q(i % n, (i + 1 + i % 3) % n, cnot)
        capfd.readouterr()
        e.flush()
        lines = [ln for ln in capfd.readouterr().err.splitlines() if ln.startswith("smem ")]
        nops = [int(ln.split("nops=")[1].split()[0]) for ln in lines]
        assert len(nops) == e.stats()["passes"] > 800 // 48
        assert max(nops) <= 48 and sum(nops) == e.stats()["micro_ops"]
        e.set_option("pass_ops", 8)                # the same program under a tighter budget
        e.reinit()
        for a in range(n):
            e.alloc_axis(a)
        for i in range(60):
            e.apply_1q(i % n, ru(2))
            e.apply_2q(i % n, (i + 1) % n, cnot)
        capfd.readouterr()
        e.flush()
        lines = [ln for ln in capfd.readouterr().err.splitlines() if ln.startswith("smem ")]
        assert lines and all(int(ln.split("nops=")[1].split()[0]) <= 16 for ln in lines)   # (a group header may ride on top)
        e.measure(0, 63, 1, False, 0.0)            # the last classical bit
        with pytest.raises(engine.EngineError, match="64 bits"):
            e.measure(1, 64, 1, False, 0.0)
        with pytest.raises(engine.EngineError, match="out of range"):
            e.cond_1q(1, np.eye(2), 64)
        e.close()
