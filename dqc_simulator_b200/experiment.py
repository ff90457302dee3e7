"""Shot-batch driver: the call a user makes to run a compiled, partitioned circuit for a batch of
noisy shots on one GPU and get per-shot results back on the host.

Mirrors ``examples/run_experiment.py:10-76`` of the reference (``take_experimental_shot`` +
fidelity against the ideal run), with the one-subprocess-per-shot loop the reference documents
(``docs/source/getting_started.rst:683-770``) replaced by one batched launch sequence.
"""
import os

import numpy as np

from . import executor, hardware, opstream

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")

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
    return opstream.load(os.path.join(DATA, f"c2_{circuit}5_3qpu_cat.json"))


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


def run_shot_batch(qpu_ops, dqc, engine, data_qubits, ideal_ket, before_flush=None, after_launch=None,
                   comm_axes_low=0, per_shot_time=True, reuse_plan=True):
    """One batch of shots: reset the register (``reinit`` also moves on to the next block of global shot
    ids, so consecutive batches are independent samples), walk the compiled op lists (enqueue only), flush
    the gate queue as fused passes, then read per-shot fidelity <phi|rho_S|phi> and the classical registers
    back to the host.  Every shot follows its own simulated timeline (``per_shot_time``).

    ``reuse_plan``: the next batch of the SAME program (same op-list and hardware objects, same options) on the
    same engine skips the executor walk and the planner: the engine replays the passes it planned for the first
    batch on the new block of shot ids (``dqe_replay``; the micro-program is uploaded again, nothing else moves)."""
    engine.reinit()
    key = (id(qpu_ops), id(dqc), comm_axes_low, per_shot_time)
    cache = getattr(engine, "_shot_batch_plan", None) if reuse_plan and hasattr(engine, "replay") else None
    if cache is not None and cache[0] == key:
        _, axes, ex = cache
        if before_flush:
            before_flush()
        engine.replay()
    else:
        if reuse_plan and hasattr(engine, "replay"):
            engine.set_option("plan_cache", 1)
        ex = executor.DQCExecutor(qpu_ops, dqc, engine, comm_axes_low=comm_axes_low, per_shot_time=per_shot_time).run()
        axes = ex.axes_of(data_qubits)
        if before_flush:
            before_flush()
        engine.flush()
        if reuse_plan and hasattr(engine, "replay"):
            engine._shot_batch_plan = (key, axes, ex)
    if after_launch:
        after_launch()
    fid = engine.fidelity_ket_all(axes, ideal_ket)
    creg = engine.read_creg_words()
    return ShotBatchResult(fid, creg, ex)


def shard_range(total_shots, rank, world):
    """Global shot ids [lo, hi) owned by `rank`: contiguous, sizes differ by at most one."""
    base, rem = divmod(int(total_shots), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def outcome_histogram(creg_words, nbits=6):
    """Counts of the low `nbits` classical bits of every shot."""
    idx = (np.asarray(creg_words, dtype=np.uint64) & np.uint64((1 << nbits) - 1)).astype(np.int64)
    return np.bincount(idx, minlength=1 << nbits).astype(np.int64)


def all_reduce_results(fidelity, hist, device=None):
    """The one collective of the shot-sharded path: sum over ranks of
    [sum of fidelities, shot count, outcome histogram].  NCCL on GPUs, gloo in CPU tests.
    Returns (mean fidelity over all shots, total shots, global histogram)."""
    import torch
    import torch.distributed as dist

    buf = torch.zeros(2 + len(hist), dtype=torch.float64, device=device or "cpu")
    buf[0] = float(np.sum(fidelity))
    buf[1] = float(len(fidelity))
    buf[2:] = torch.as_tensor(np.asarray(hist, dtype=np.float64)).to(buf.device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(buf)
    out = buf.cpu().numpy()
    return float(out[0] / max(out[1], 1.0)), int(round(out[1])), np.rint(out[2:]).astype(np.int64)


def fidelity_squared(rho, reference):
    """``qapi.fidelity(qubits, reference, squared=True)`` on a reduced density matrix read back from
    the engine (``util/helper.py:151, 209-214``; ``examples/run_experiment.py:74-75``): <phi|rho|phi> for
    a ket reference, Uhlmann's (Tr sqrt(sqrt(sigma) rho sqrt(sigma)))^2 for a density-matrix
    reference.  Host-side, for the small (<= 2^5) matrices the read-out produces."""
    rho = np.asarray(rho, dtype=complex)
    ref = np.asarray(reference, dtype=complex)
    if ref.ndim == 1 or 1 in ref.shape:
        v = ref.reshape(-1)
        return float(np.real(v.conj() @ rho @ v))
    lam, vec = np.linalg.eigh(ref)
    sq = (vec * np.sqrt(np.clip(lam, 0, None))) @ vec.conj().T
    ev = np.linalg.eigvalsh(sq @ rho @ sq)
    return float(np.sum(np.sqrt(np.clip(ev, 0, None))) ** 2)


class DataCollector:
    """Result frame of a finished run, in the shape of the reference's
    ``netsquid.util.datacollector.DataCollector`` as ``util/helper.py:112-163`` configures it
    (``get_data_collector`` / ``get_data_collector4dm``): ``dc.dataframe`` is a pandas frame with NetSquid's
    ``time_stamp`` (simulated ns at which ``DQCMasterProtocol`` signalled FINISHED) and ``entity_name``
    columns plus ``fidelity`` -- one row per shot of the batch (the reference produces one row per process,
    ``docs/source/getting_started.rst:683-770``), so ``dc.dataframe["fidelity"].iloc[0]`` reads the same."""

    def __init__(self, executor, qubit_indices_2b_checked, desired_state):
        self.executor = executor
        self.qubits = []
        for tup in qubit_indices_2b_checked:
            first, node = tup[0], tup[-1]
            if isinstance(first, (int, np.integer)):
                idx = [int(first)]
            elif isinstance(first, (list, tuple)):
                idx = [int(i) for i in first]
            else:   # same check and wording as util/helper.py:139-146
                raise TypeError(f"{tup} not of correct form."
                                "The first element of each tuple must be qubit"
                                " index or list of qubit indices")
            name = node if isinstance(node, str) else node.name
            self.qubits += [(i, name) for i in idx]
        self.desired_state = np.asarray(desired_state, dtype=complex)
        self._frame = None

    def collect(self):
        """``qmemory.pop`` of the listed qubits (memory noise up to the end of the run), then
        ``qapi.fidelity(qubits, desired_state, squared=True)`` per shot; the qubits are discarded."""
        import pandas as pd

        ex, be = self.executor, self.executor.backend
        axes = ex.axes_of(self.qubits)
        ref = self.desired_state
        if ref.ndim == 1 or 1 in ref.shape:
            ket = ref.reshape(-1)
            if hasattr(be, "fidelity_ket_all"):   # the CUDA engine: one kernel for the whole batch
                fid = np.asarray(be.fidelity_ket_all(axes, ket), dtype=float)
            else:
                fid = np.array([be.fidelity_ket(axes, ket, s) for s in range(be.batch)], dtype=float)
        else:   # density-matrix reference: Uhlmann on the host, one small reduced matrix per shot
            fid = np.array([fidelity_squared(be.reduced_dm(axes, s), ref) for s in range(be.batch)])
        # simulated ns at which the run finished: one value per shot once the timeline depends on outcomes
        end = ex.read_time(ex.end_time) if ex.T.is_reg(ex.end_time) else float(ex.end_time)
        self._frame = pd.DataFrame({"time_stamp": end, "entity_name": "DQCMasterProtocol", "fidelity": fid})
        return self._frame

    @property
    def dataframe(self):
        return self._frame if self._frame is not None else self.collect()


def get_data_collector(executor, qubit_indices_2b_checked, desired_state):
    """Same call shape as ``util/helper.py:112`` with the executor in the place of the master protocol:
    ``qubit_indices_2b_checked`` = [(index or [indices], node or node name), ...], ``desired_state`` a ket or
    a density matrix.  Read ``.dataframe`` after ``executor.run()``."""
    return DataCollector(executor, qubit_indices_2b_checked, desired_state)


get_data_collector4dm = get_data_collector   # util/helper.py:166-223: same frame, density-matrix reference


class MidSimDataCollector:
    """Results of the instructions a circuit marks ``"logged"`` (``dqc_control.py:754-791`` signals
    ``QDCSignals.RESULT_PRODUCED`` with ``(result, ancilla_qubit_index)``), in the frame of
    ``util/helper.py:236-263`` (``get_data_collector_for_mid_sim_instr_output``): NetSquid's ``time_stamp`` /
    ``entity_name`` columns plus ``result`` and ``ancilla_qubit_index`` -- one row per logged instruction and shot
    (column ``shot``; the reference runs one shot per process and so has one row per logged instruction)."""

    def __init__(self, executor):
        self.executor = executor
        self._frame = None

    def collect(self):
        import pandas as pd

        ex, be = self.executor, self.executor.backend
        rows = []
        for (node, index, res), when in zip(ex.logged_results, ex.logged_times):
            if hasattr(res, "index"):          # outcome held per shot on the device
                vals = np.asarray(be.read_creg(res.index), dtype=np.int64)
            elif res is None:
                vals = np.full(be.batch, -1, dtype=np.int64)
            else:
                vals = np.full(be.batch, int(res), dtype=np.int64)
            t = ex.read_time(when) if ex.T.is_reg(when) else np.full(be.batch, float(when))
            idx = index[0] if isinstance(index, (list, tuple)) and len(index) == 1 else index
            rows.append(pd.DataFrame({"time_stamp": np.asarray(t, dtype=float), "entity_name": f"interpreter_{node}",
                                      "result": vals, "ancilla_qubit_index": [idx] * be.batch,
                                      "shot": np.arange(be.batch)}))
        cols = ["time_stamp", "entity_name", "result", "ancilla_qubit_index", "shot"]
        self._frame = pd.concat(rows, ignore_index=True) if rows else pd.DataFrame(columns=cols)
        return self._frame

    @property
    def dataframe(self):
        return self._frame if self._frame is not None else self.collect()


def get_data_collector_for_mid_sim_instr_output(executor):
    """``util/helper.py:236``: collector of the mid-simulation instruction outputs; read ``.dataframe`` after
    ``executor.run()``."""
    return MidSimDataCollector(executor)
