"""Live front-end path: a ``.qasm`` file and a hardware spec in, the circuit run on the engine out.

The reference's front end (``software/qasm2ast.py`` -> ``ast2dqc_circuit`` -> ``compiler_preprocessing`` ->
``partitioner`` -> ``compilers``) is pure Python and runs UNMODIFIED, in-process, from wherever the
``dqc_simulator`` package is importable (an installed package, ``$DQC_SIMULATOR_SRC``, or the ``src`` argument);
where no NetSquid is installed the name shim of ``compat/netsquid_shim.py`` stands in for the symbols it
imports.  Mirrors ``examples/setup_sim.py:16-43`` (first-come-first-served allocation, ``partition_gate_tuples``,
``sort_many_qpus_greedily_by_node_and_time``) minus the DES objects, then hands the ``qpu_op_dict`` to
``executor.DQCExecutor`` and the C-ABI."""
import importlib
import os
import sys
import warnings


class ReferenceUnavailable(RuntimeError):
    pass


def _import_reference(src=None):
    """Make ``dqc_simulator`` importable; returns the modules of its front end."""
    try:
        import netsquid  # noqa: F401  (a real install, Python 3.9)
    except ModuleNotFoundError:
        from . import netsquid_shim

        netsquid_shim.install()
    for cand in (src, os.environ.get("DQC_SIMULATOR_SRC"), "/root/reference/src"):
        try:
            importlib.import_module("dqc_simulator")
            break
        except ModuleNotFoundError:
            if cand and os.path.isdir(cand) and cand not in sys.path:
                sys.path.insert(0, cand)
    try:
        warnings.simplefilter("ignore")
        pre = importlib.import_module("dqc_simulator.software.compiler_preprocessing")
        comp = importlib.import_module("dqc_simulator.software.compilers")
        part = importlib.import_module("dqc_simulator.software.partitioner")
    except ModuleNotFoundError as e:
        raise ReferenceUnavailable(f"the dqc_simulator package is not importable here ({e}); set DQC_SIMULATOR_SRC "
                                   "to its src/ directory, or use the pre-compiled op streams (opstream.load)") from e
    return pre, comp, part


def reference_available(src=None):
    try:
        _import_reference(src)
        return True
    except ReferenceUnavailable:
        return False


class _Qmem:
    def __init__(self, num_positions, num_comm):
        self.comm_qubit_positions = list(range(num_comm))
        self.processing_qubit_positions = list(range(num_comm, num_positions))


class _Node:
    """The two attributes ``first_come_first_served_qubits_to_qpus`` reads (``partitioner.py:117-146``)."""

    def __init__(self, i, num_positions, num_comm):
        self.name = f"node_{i}"
        self.qmemory = _Qmem(num_positions, num_comm)


def compile_qasm(qasm_path, num_qpus, num_positions, num_comm_qubits=2, scheme="cat", include_path=None, src=None):
    """QASM file -> (qpu_op_dict, meta) through the reference front end.  ``num_qpus == 1``: monolithic (every gate
    local); otherwise first-come-first-served allocation over the QPUs and remote gates by ``scheme``
    (``cat`` / ``1tp`` / ``tp_safe``), exactly ``examples/setup_sim.py:16-43``."""
    pre, comp, part = _import_reference(src)
    include_path = include_path or os.path.dirname(os.path.abspath(qasm_path))
    mono = pre.preprocess_qasm_to_compilable_monolithic(qasm_path, include_path=include_path).ops
    if num_qpus == 1:
        parts = [tuple("node_0" if e == "monolithic_qc" else e for e in g) for g in mono]
        ops = comp.sort_greedily_by_node_and_time(parts)
        n = len(mono[0][1])
        return ops, {"num_qpus": 1, "num_positions": num_positions, "num_comm_qubits": 0, "monolithic_ops": len(mono),
                     "qubit_map": {str(i): [i, "node_0"] for i in range(n)}}
    nodes = [_Node(i, num_positions, num_comm_qubits) for i in range(num_qpus)]
    lookup, alloc = part.first_come_first_served_qubits_to_qpus(mono, nodes)
    parts = part.partition_gate_tuples(mono, None, scheme, lookup, alloc)
    sorter = comp.sort_many_qpus_greedily_by_node_and_time if num_qpus > 2 else comp.sort_greedily_by_node_and_time
    ops = sorter(parts)
    return ops, {"num_qpus": num_qpus, "num_positions": num_positions, "num_comm_qubits": num_comm_qubits,
                 "scheme": scheme, "compiler": sorter.__name__, "monolithic_ops": len(mono),
                 "qubit_map": {str(k): [v[0], v[1]] for k, v in lookup.items()}, "proc_alloc": alloc}


def run_qasm(qasm_path, dqc_spec, scheme="cat", engine=None, formalism="ket", batch=1, seed=0, max_qubits=None,
             include_path=None, src=None, **executor_kw):
    """Run a QASM circuit on the hardware ``dqc_spec`` (``hardware.DQCSpec``): reference front end -> executor ->
    engine.  ``engine`` defaults to a fresh CUDA ``Engine`` sized for the data qubits plus one live ebit pair per
    QPU pair.  Returns (executor, engine, meta); read states with ``engine.reduced_dm(executor.axes_of(...))``."""
    from .. import executor
    from ..engine import Engine

    nq = len(dqc_spec.qpus)
    spec0 = dqc_spec.qpus[0]
    ops, meta = compile_qasm(qasm_path, nq, spec0.num_positions, getattr(spec0, "num_comm_qubits", 2) if nq > 1 else 0,
                             scheme=scheme, include_path=include_path, src=src)
    if engine is None:
        n_data = len(meta["qubit_map"])
        engine = Engine(formalism, max_qubits or (n_data + (2 if nq > 1 else 0)), batch, seed=seed)
    ex = executor.DQCExecutor(ops, dqc_spec, engine, **executor_kw).run()
    return ex, engine, meta
