"""Generate the committed op-stream fixtures by running the REFERENCE front end, unmodified.

Run in the build container only (needs ``/root/reference``):

    python tests/golden/make_golden.py

The reference's pure-Python front end (qasm2ast -> ast2dqc_circuit -> partitioner ->
compilers) is imported from ``/root/reference/src`` behind the NetSquid name stub in
``tools/ref_frontend/netsquid_stub.py`` (NetSquid itself cannot be installed here,
SURVEY.md section 8c) and its compiled ``qpu_op_dict`` outputs are serialised with
``dqc_simulator_b200.opstream``.  The GPU box has no ``/root/reference``; tests and
``bench.py`` read only the fixtures written here.

Fixtures
--------
* ``c1_ghz5_2qpu_bisect_cat``     BASELINE config 1 (``preprocess_qasm_to_compilable_bipartitioned``)
* ``c2_{ghz,grover,qft}5_3qpu_cat``  config 2, exactly ``examples/setup_sim.py:16-43`` on the
  hardware of ``examples/setup_hardware.py:11-55`` (10 positions, 2 comm qubits, 3 QPUs)
* ``c3_qft{N}_4qpu_cat``          config 3 shape (N = 8, 12 for oracle-sized parity, 26 full)
* ``c4_qft{N}_mono``              config 4 shape: one QPU, all gates local (N = 6, 16; 17 = 256 GiB, one rho sharded over 8 GPUs)
* ``c5_{ghz,qft}{N}_mono``        config 5 shape (N = 12 parity size, 30 .. 34 full: weak scaling over 1, 2, 4, 8, 16 GPUs)
* ``mqt/{name}_{mono,3qpu_cat,3qpu_tp_safe}``  the 22 MQT-bench circuits of ``tests/MQT_benchmarking_circuits`` (the
  reference's front-end test inputs), monolithic and partitioned over 3 QPUs
* ``ref_*``                       gate-tuple circuits lifted from the reference's own tests
  (``tests/unit_tests/software/test_dqc_control.py:62-162, 249-306``,
  ``tests/unit_tests/qlib/test_circuit_identities.py``) compiled by the reference compiler.
"""
import os
import sys
import tempfile
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools", "ref_frontend"))

import numpy as np  # noqa: E402

import netsquid_stub  # noqa: E402
import qasm_textbook  # noqa: E402

netsquid_stub.install()
warnings.simplefilter("ignore")

from netsquid.components import instructions as instr  # noqa: E402  (stub)
from dqc_simulator.software.compiler_preprocessing import (  # noqa: E402
    preprocess_qasm_to_compilable_bipartitioned,
    preprocess_qasm_to_compilable_monolithic,
)
from dqc_simulator.software.compilers import (  # noqa: E402
    sort_greedily_by_node_and_time,
    sort_many_qpus_greedily_by_node_and_time,
)
from dqc_simulator.software.partitioner import (  # noqa: E402
    first_come_first_served_qubits_to_qpus,
    partition_gate_tuples,
)
from dqc_simulator.qlib import circuit_identities  # noqa: E402

from dqc_simulator_b200 import circuits, opstream  # noqa: E402

EX = "/root/reference/examples"


class _FakeQmem:
    def __init__(self, num_positions, num_comm):
        self.comm_qubit_positions = list(range(num_comm))
        self.processing_qubit_positions = list(range(num_comm, num_positions))


class _FakeNode:
    """Carries the two attributes ``first_come_first_served_qubits_to_qpus`` reads
    (``partitioner.py:117-146``): ``name`` and ``qmemory.processing_qubit_positions``."""

    def __init__(self, i, num_positions, num_comm):
        self.name = f"node_{i}"
        self.qmemory = _FakeQmem(num_positions, num_comm)


def _qasm_path(text):
    f = tempfile.NamedTemporaryFile("w", suffix=".qasm", delete=False)
    f.write(text)
    f.close()
    return f.name


def compile_fcfs(qasm_file, num_qpus, num_positions, num_comm, scheme="cat"):
    """examples/setup_sim.py:16-43 without the DES objects."""
    nodes = [_FakeNode(i, num_positions, num_comm) for i in range(num_qpus)]
    mono = preprocess_qasm_to_compilable_monolithic(qasm_file, include_path=EX).ops
    lookup, alloc = first_come_first_served_qubits_to_qpus(mono, nodes)
    parts = partition_gate_tuples(mono, None, scheme, lookup, alloc)
    comp = sort_many_qpus_greedily_by_node_and_time if num_qpus > 2 else sort_greedily_by_node_and_time
    ops = comp(parts)
    meta = {
        "num_qpus": num_qpus, "num_positions": num_positions, "num_comm_qubits": num_comm,
        "scheme": scheme, "compiler": comp.__name__, "monolithic_ops": len(mono),
        "qubit_map": {str(k): [v[0], v[1]] for k, v in lookup.items()},
        "proc_alloc": alloc,
    }
    return ops, meta


def compile_mono(qasm_file, num_positions):
    mono = preprocess_qasm_to_compilable_monolithic(qasm_file, include_path=EX).ops
    parts = []
    for g in mono:
        g = list(g)
        # monolithic gate tuples name a single processor 'monolithic_qc'
        parts.append(tuple("node_0" if e == "monolithic_qc" else e for e in g))
    ops = sort_greedily_by_node_and_time(parts)
    n = len(mono[0][1])
    meta = {"num_qpus": 1, "num_positions": num_positions, "num_comm_qubits": 0,
            "monolithic_ops": len(mono),
            "qubit_map": {str(i): [i, "node_0"] for i in range(n)}}
    return ops, meta


def save(name, ops, meta, gz=False):
    path = os.path.join(HERE, name + (".json.gz" if gz else ".json"))
    opstream.save(path, ops, meta)
    n_ops = sum(len(sl) for s in ops.values() for sl in s)
    print(f"{os.path.basename(path):40s} nodes={len(ops)} slices={max(len(s) for s in ops.values())} ops={n_ops}")


def main():
    # ---- config 1 ----------------------------------------------------------------------
    c = preprocess_qasm_to_compilable_bipartitioned(f"{EX}/ghz_5qubits.qasm", "cat", include_path=EX)
    ops = sort_greedily_by_node_and_time(c.ops)
    save("c1_ghz5_2qpu_bisect_cat", ops, {
        "num_qpus": 2, "num_comm_qubits": 2, "node_sizes": c.node_sizes, "scheme": "cat",
        "num_positions": max(c.node_sizes.values()),
        "qubit_map": {"0": [2, "node_0"], "1": [3, "node_0"], "2": [4, "node_0"],
                      "3": [2, "node_1"], "4": [3, "node_1"]}})
    # ---- config 2 ----------------------------------------------------------------------
    for circ in ("ghz", "grover", "qft"):
        ops, meta = compile_fcfs(f"{EX}/{circ}_5qubits.qasm", 3, 10, 2)
        save(f"c2_{circ}5_3qpu_cat", ops, meta)
    # ---- config 3 ----------------------------------------------------------------------
    for n in (8, 12, 26):
        ops, meta = compile_fcfs(_qasm_path(circuits.qft_qasm(n)), 4, 2 + (n + 3) // 4, 2)
        save(f"c3_qft{n}_4qpu_cat", ops, meta, gz=n > 12)
    for n in (10,):
        ops, meta = compile_fcfs(_qasm_path(circuits.ghz_qasm(n)), 4, 2 + (n + 3) // 4, 2)
        save(f"c3_ghz{n}_4qpu_cat", ops, meta)
    # ---- config 4 / 5 shapes (single register, all local) ---------------------------------
    for n in (6, 16, 17):
        ops, meta = compile_mono(_qasm_path(circuits.qft_qasm(n)), n)
        save(f"c4_qft{n}_mono", ops, meta, gz=n > 12)
    for kind, gen in (("ghz", circuits.ghz_qasm), ("qft", circuits.qft_qasm)):
        for n in (12, 30, 31, 32, 33, 34):
            ops, meta = compile_mono(_qasm_path(gen(n)), n)
            save(f"c5_{kind}{n}_mono", ops, meta, gz=n > 12)
    # ---- the reference's MQT-bench test circuits (tests/MQT_benchmarking_circuits, 22 x 5 qubits):
    #      every gate family its front end lowers (u1/u2/u3, rz, sx, cp, ccx decompositions, swap, ...),
    #      once on a single QPU and once first-come-first-served over 3 QPUs with cat-comm remote gates
    mqt_dir = "/root/reference/tests/MQT_benchmarking_circuits"
    os.makedirs(os.path.join(HERE, "mqt"), exist_ok=True)
    for f in sorted(os.listdir(mqt_dir)):
        if not f.endswith(".qasm"):
            continue
        stem = f[:-len("_indep_qiskit_5.qasm")]
        # expected output state: an independent simulation of the QASM text (tools/ref_frontend/
        # qasm_textbook.py) with the reference's two macro deviations replayed; plus how far plain qelib1.inc
        # semantics are from it (1.0 unless the circuit uses cu1 / cp / crx)
        text = open(os.path.join(mqt_dir, f)).read()
        want = qasm_textbook.simulate(text, reference_macros=True)
        plain = qasm_textbook.simulate(text, reference_macros=False)
        extra = {"source": f"tests/MQT_benchmarking_circuits/{f}",
                 "expected_ket": [[float(z.real), float(z.imag)] for z in want],
                 "qelib1_fidelity": float(abs(np.vdot(plain, want)) ** 2)}
        ops, meta = compile_mono(os.path.join(mqt_dir, f), 5)
        save(f"mqt/{stem}_mono", ops, dict(meta, **extra), gz=True)
        ops, meta = compile_fcfs(os.path.join(mqt_dir, f), 3, 10, 2)
        save(f"mqt/{stem}_3qpu_cat", ops, dict(meta, **extra), gz=True)
        # the same partition with teleportation-based remote gates (scheme 'tp_safe': teleport the control
        # over, apply the gate locally, teleport it back -- compilers.py:255-278)
        ops, meta = compile_fcfs(os.path.join(mqt_dir, f), 3, 10, 2, scheme="tp_safe")
        save(f"mqt/{stem}_3qpu_tp_safe", ops, dict(meta, **extra), gz=True)
    # ---- circuits from the reference's own tests ---------------------------------------------
    I = instr
    ref = {
        # test_dqc_control.py:62-80
        "ref_local_x": [(I.INSTR_INIT, 2, "node_0"), (I.INSTR_INIT, 2, "node_1"),
                        (I.INSTR_X, 2, "node_0"), (I.INSTR_X, 2, "node_1")],
        # :82-96
        "ref_remote_cnot_cat": [(I.INSTR_INIT, 2, "node_0"), (I.INSTR_INIT, 2, "node_1"),
                                (I.INSTR_H, 2, "node_0"),
                                (I.INSTR_CNOT, 2, "node_0", 2, "node_1", "cat")],
        # :98-110
        "ref_teleport_only": [(I.INSTR_INIT, 2, "node_0"), (I.INSTR_H, 2, "node_0"),
                              ([], 2, "node_0", 0, "node_1", "teleport_only")],
        # :112-126
        "ref_remote_cnot_tp_safe": [(I.INSTR_INIT, 2, "node_0"), (I.INSTR_INIT, 2, "node_1"),
                                    (I.INSTR_H, 2, "node_0"),
                                    (I.INSTR_CNOT, 2, "node_0", 2, "node_1", "tp_safe")],
        # :128-144
        "ref_3x_remote_cnot_cat": [(I.INSTR_INIT, 2, "node_0"), (I.INSTR_INIT, 2, "node_1"),
                                   (I.INSTR_H, 2, "node_0")]
        + [(I.INSTR_CNOT, 2, "node_0", 2, "node_1", "cat")] * 3,
        # :146-162
        "ref_3x_remote_cnot_tp_safe": [(I.INSTR_INIT, 2, "node_0"), (I.INSTR_INIT, 2, "node_1"),
                                       (I.INSTR_H, 2, "node_0")]
        + [(I.INSTR_CNOT, 2, "node_0", 2, "node_1", "tp_safe")] * 3,
        # 1tp variant (compilers.py:183-195)
        "ref_remote_cnot_1tp": [(I.INSTR_INIT, 2, "node_0"), (I.INSTR_INIT, 2, "node_1"),
                                (I.INSTR_H, 2, "node_0"),
                                (I.INSTR_CNOT, 2, "node_0", 2, "node_1", "1tp")],
        # test_dqc_control.py:249-271 (logged measurements -> [1, 0, 1])
        "ref_logged_meas": [(I.INSTR_INIT, [2, 3, 4], "node_0"), (I.INSTR_X, 2, "node_0"),
                            (I.INSTR_X, 4, "node_0"),
                            (I.INSTR_MEASURE, 2, "node_0", "logged"),
                            (I.INSTR_MEASURE, 3, "node_0", "logged"),
                            (I.INSTR_MEASURE, 4, "node_0", "logged")],
        # :273-306
        "ref_distribute_ebit": [("node_0", "node_1", "distribute_ebit")],
        "ref_distribute_3_ebits": [("node_0", "node_1", "distribute_ebit")] * 3,
        # examples/getting_started.py:76-85 (3 QPUs, remote CNOT Bell pair)
        "ref_getting_started": [(I.INSTR_INIT, [2, 3, 4], "node_0"),
                                (I.INSTR_INIT, [2, 3, 4], "node_1"),
                                (I.INSTR_INIT, [2, 3, 4], "node_2"),
                                (I.INSTR_H, 2, "node_0"),
                                (I.INSTR_CNOT, 2, "node_0", 2, "node_1", "cat")],
    }
    for name, tuples in ref.items():
        n_nodes = len({e for t in tuples for e in t if isinstance(e, str) and e.startswith("node_")})
        comp = sort_many_qpus_greedily_by_node_and_time if n_nodes > 2 else sort_greedily_by_node_and_time
        save(name, comp(tuples), {"source": "reference tests / examples", "compiler": comp.__name__})
    # Toffoli decompositions of qlib/circuit_identities.py:15-122, local and distributed (cat)
    for a, b, t in ((0, 0, 0), (0, 1, 0), (1, 0, 0), (1, 1, 0), (1, 1, 1)):
        pre = [(I.INSTR_INIT, [2, 3, 4], "node_0"), (I.INSTR_INIT, [2], "node_1")]
        if a:
            pre.append((I.INSTR_X, 2, "node_0"))
        if b:
            pre.append((I.INSTR_X, 3, "node_0"))
        if t:
            pre.append((I.INSTR_X, 4, "node_0"))
        tof = circuit_identities.two_control_ibm_toffoli_decomp(2, "node_0", 3, "node_0", 4, "node_0")
        save(f"ref_toffoli_local_{a}{b}{t}", sort_greedily_by_node_and_time(pre + tof),
             {"source": "qlib/circuit_identities.py:15-122", "inputs": [a, b, t]})
        pre = [(I.INSTR_INIT, [2, 3], "node_0"), (I.INSTR_INIT, [2], "node_1")]
        if a:
            pre.append((I.INSTR_X, 2, "node_0"))
        if b:
            pre.append((I.INSTR_X, 3, "node_0"))
        if t:
            pre.append((I.INSTR_X, 2, "node_1"))
        tof = circuit_identities.two_control_ibm_toffoli_decomp(2, "node_0", 3, "node_0", 2, "node_1",
                                                            scheme="cat")
        save(f"ref_toffoli_remote_cat_{a}{b}{t}", sort_greedily_by_node_and_time(pre + tof),
             {"source": "qlib/circuit_identities.py:15-122", "inputs": [a, b, t]})


if __name__ == "__main__":
    main()
