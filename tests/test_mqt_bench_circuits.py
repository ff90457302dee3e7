"""The reference's own front-end test inputs -- the 22 five-qubit MQT-bench circuits of
``tests/MQT_benchmarking_circuits`` (ae, dj, ghz, graphstate, grover x2, qaoa, qft, qpe, qwalk, random, vqe,
wstate, ...) -- run end to end: reference front end (fixtures: ``tests/golden/mqt``, made by
``tests/golden/make_golden.py``) -> executor -> state engine.

Expected states come from an INDEPENDENT simulation of the QASM text (``tools/ref_frontend/
qasm_textbook.py``, qelib1.inc gate definitions, no reference code) with the reference's two documented macro
deviations replayed (cu1 / cp / crx); ``meta["qelib1_fidelity"]`` records how far plain
OpenQASM semantics are from what the reference computes (exactly 1 for the 13 circuits without those gates).
Covered (the tp_safe partition on the oracle only so far): every gate family the front end lowers (u1/u2/u3, p, rx/ry/rz, h, x, z, t, tdg, cx, cz, cp, cu1, cu3,
crx, ch, rzz, swap, ccx, rccx), the single-QPU path, and the 3-QPU partition with cat-comm remote gates and up
to 540 sampled mid-circuit measurements per circuit."""
import glob
import os

import numpy as np
import pytest

from dqc_simulator_b200 import executor, hardware, opstream
from oracle.numpy_engine import OracleEngine

MQT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mqt")
NAMES = sorted(os.path.basename(f)[:-len("_mono.json.gz")] for f in glob.glob(os.path.join(MQT, "*_mono.json.gz")))
NOISE = dict(p_depolar_error_cnot=1e-3, single_qubit_gate_error_prob=2e-5, meas_error_prob=3e-3,
             comm_qubit_depolar_rate=0.055, proc_qubit_depolar_rate=0.055)
# one circuit per gate family for the (slower) noisy density-matrix engine-vs-oracle comparison
NOISY_SUBSET = ["ae", "qpeexact", "random", "qwalk-v-chain", "portfolioqaoa", "qnn", "wstate", "graphstate"]


def _load(name, kind):
    ops, meta = opstream.load(os.path.join(MQT, f"{name}_{kind}.json.gz"))
    want = np.array([complex(re, im) for re, im in meta["expected_ket"]])
    qubits = [tuple(meta["qubit_map"][str(i)]) for i in range(5)]
    return ops, meta, want, qubits


def _run(make_backend, name, kind, formalism, batch, seed, **noise):
    ops, meta, want, qubits = _load(name, kind)
    nq = 1 if kind == "mono" else 3
    if kind == "mono":
        dqc = hardware.DQCSpec.uniform(1, num_positions=5, num_comm_qubits=0, **noise)
        be = make_backend(formalism, 5, batch, seed)
    else:
        state = hardware.werner_state(0.99) if noise else None
        dqc = hardware.DQCSpec.uniform(nq, state4distribution=state, num_positions=10, num_comm_qubits=2, **noise)
        be = make_backend(formalism, 7, batch, seed)
    ex = executor.DQCExecutor(ops, dqc, be).run()
    return ex, be, ex.axes_of(qubits), want


def cpu(formalism, n, batch, seed):
    return OracleEngine(formalism, n, batch, seed=seed)


def gpu(formalism, n, batch, seed):
    from dqc_simulator_b200.engine import Engine

    return Engine(formalism, n, batch, seed=seed)


def test_all_22_reference_test_circuits_are_present():
    assert len(NAMES) == 22
    plain = {n: _load(n, "mono")[1]["qelib1_fidelity"] for n in NAMES}
    assert sum(abs(v - 1) < 1e-9 for v in plain.values()) == 13   # the rest use the reference's cu1 / cp / crx macros


@pytest.mark.parametrize("name", NAMES)
def test_oracle_monolithic_matches_independent_simulation(name):
    ex, be, axes, want = _run(cpu, name, "mono", "ket", 1, 1)
    rho = be.reduced_dm(axes, 0)
    assert abs(np.real(want.conj() @ rho @ want) - 1) < 1e-10


@pytest.mark.parametrize("name", NAMES)
def test_oracle_partitioned_over_3_qpus_matches_independent_simulation(name):
    """cat-comm remote gates with sampled measurement outcomes and their classically controlled corrections:
    the reduced state of the data qubits is the monolithic output state for every outcome history."""
    ex, be, axes, want = _run(cpu, name, "3qpu_cat", "ket", 2, 7)
    for shot in range(2):
        rho = be.reduced_dm(axes, shot)
        assert abs(np.real(want.conj() @ rho @ want) - 1) < 1e-10


@pytest.mark.parametrize("name", NAMES)
def test_oracle_partitioned_with_teleportation_matches_independent_simulation(name):
    """Same partition, remote gates by scheme 'tp_safe' (teleport the control qubit over, gate locally,
    teleport back; ``compilers.py:255-278``, ``dqc_control.py:405-613``): Bell-state measurements, both
    corrections, the comm-qubit swap -- up to 1080 sampled measurements per circuit (qwalk)."""
    ops, meta, want, qubits = _load(name, "3qpu_tp_safe")
    be = cpu("ket", 9, 2, 3)
    ex = executor.DQCExecutor(ops, hardware.DQCSpec.uniform(3, num_positions=10, num_comm_qubits=2), be).run()
    axes = ex.axes_of(qubits)
    for shot in range(2):
        rho = be.reduced_dm(axes, shot)
        assert abs(np.real(want.conj() @ rho @ want) - 1) < 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_engine_partitioned_matches_independent_simulation(name):
    ex, be, axes, want = _run(gpu, name, "3qpu_cat", "ket", 3, 11)
    fid = be.fidelity_ket_all(axes, want)
    assert np.abs(fid - 1).max() < 1e-10
    ex, be, axes, want = _run(gpu, name, "mono", "ket", 1, 11)
    assert abs(be.fidelity_ket(axes, want, 0) - 1) < 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize("name", NOISY_SUBSET)
def test_cuda_engine_noisy_dm_matches_oracle(name):
    """Noisy hardware of examples/run_experiment.py:80-87 on the partitioned circuit, density matrices, two
    sampled trajectories: rho of the engine = rho of the oracle (1e-10), classical registers bit-exact."""
    res = []
    for mk in (gpu, cpu):
        ex, be, axes, _ = _run(mk, name, "3qpu_cat", "dm", 2, 5, **NOISE)
        res.append(([be.reduced_dm(axes, s) for s in range(2)], np.asarray(be.read_creg_words() if mk is gpu else be.creg)))
    for a, b in zip(res[0][0], res[1][0]):
        assert np.abs(a - b).max() < 1e-10
    assert np.array_equal(res[0][1].astype(np.uint64), res[1][1].astype(np.uint64))
