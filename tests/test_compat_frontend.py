"""Live front-end path: ``compat.frontend.run_qasm`` runs the reference's UNMODIFIED front end
in-process (qasm2ast -> ast2dqc_circuit -> partitioner -> compilers, behind ``compat.netsquid_shim``) and hands the
result to the executor -- no fixture in between.  Needs the reference package importable (this build container:
``/root/reference/src``); skipped elsewhere (the GPU box holds only the serialised op streams)."""
import glob
import os

import numpy as np
import pytest

from dqc_simulator_b200 import hardware, opstream
from dqc_simulator_b200.compat import frontend

HERE = os.path.dirname(os.path.abspath(__file__))
MQT_QASM = "/root/reference/tests/MQT_benchmarking_circuits"
pytestmark = pytest.mark.skipif(not (os.path.isdir(MQT_QASM) and frontend.reference_available()),
                                reason="the reference package is not importable here")
FILES = sorted(glob.glob(os.path.join(MQT_QASM, "*.qasm")))


def _fixture(qasm, kind):
    stem = os.path.basename(qasm)[:-len("_indep_qiskit_5.qasm")]
    path = os.path.join(HERE, "golden", "mqt", f"{stem}_{kind}.json.gz")
    assert os.path.exists(path), path
    return path


def test_the_reference_test_inputs_are_all_there():
    assert len(FILES) == 22


@pytest.mark.parametrize("qasm", FILES, ids=[os.path.basename(f)[:-5] for f in FILES])
def test_live_front_end_equals_committed_op_streams_and_runs(qasm):
    """(1) What the front end emits live is, op for op and matrix for matrix, what the committed fixtures hold
    (monolithic and 3 QPUs / cat-comm); (2) ``run_qasm`` on the live path ends in the circuit's output state
    (the fixture's independently simulated ``expected_ket``) for sampled cat-comm outcomes."""
    from oracle.numpy_engine import OracleEngine

    for kind, nq, npos, ncomm in (("mono", 1, 5, 0), ("3qpu_cat", 3, 10, 2)):
        ops, meta = frontend.compile_qasm(qasm, nq, npos, ncomm, scheme="cat", include_path="/root/reference/examples")
        fx = _fixture(qasm, kind)
        want_ops, want_meta = opstream.load(fx)
        live, held = opstream.encode(ops)["qpu_ops"], opstream.encode(want_ops)["qpu_ops"]
        assert live == held, f"{os.path.basename(qasm)} ({kind}): live front end differs from the fixture"
        assert meta["qubit_map"] == want_meta["qubit_map"]
    dqc = hardware.DQCSpec.uniform(3, num_positions=10, num_comm_qubits=2)
    ex, be, meta = frontend.run_qasm(qasm, dqc, scheme="cat", engine=OracleEngine("ket", 7, 2, seed=3),
                                     include_path="/root/reference/examples")
    want = np.array([complex(re, im) for re, im in opstream.load(_fixture(qasm, "mono"))[1]["expected_ket"]])
    qubits = [tuple(meta["qubit_map"][str(i)]) for i in range(5)]
    for shot in range(2):
        rho = be.reduced_dm(ex.axes_of(qubits), shot)
        assert abs(np.real(want.conj() @ rho @ want) - 1) < 1e-10


def test_product_copies_of_the_benchmark_circuits_match_the_fixtures():
    for c in ("ghz", "grover", "qft"):
        a = open(os.path.join(HERE, "golden", f"c2_{c}5_3qpu_cat.json")).read()
        b = open(os.path.join(HERE, "..", "dqc_simulator_b200", "data", f"c2_{c}5_3qpu_cat.json")).read()
        assert a == b
