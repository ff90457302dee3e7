"""CPU suite: Philox known answers, op-stream serialisation, synthetic circuits vs the reference's
shipped QASM, and the C-ABI library loading with every declared symbol (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest

from dqc_simulator_b200 import circuits, engine, opstream
from oracle import philox

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_philox_known_answers():
    """Random123 kat_vectors: philox4x32-10."""
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
        ((0xFFFFFFFF,) * 4, (0xFFFFFFFF, 0xFFFFFFFF), (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
        ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
         (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
    ]
    for ctr, key, want in kat:
        got = philox.philox4x32_10(np.array(ctr, dtype=np.uint32), np.array(key, dtype=np.uint32))
        assert tuple(int(x) for x in got) == want


def test_uniforms_in_unit_interval_and_keyed_by_global_shot():
    u, u2 = philox.shot_uniforms(2024, np.arange(1000), 3)
    assert (u >= 0).all() and (u < 1).all() and (u2 >= 0).all() and (u2 < 1).all()
    a, _ = philox.shot_uniforms(2024, np.arange(500, 1000), 3)
    assert np.array_equal(u[500:], a)
    assert abs(u.mean() - 0.5) < 0.05


def test_opstream_roundtrip(golden_dir):
    for name in ("c1_ghz5_2qpu_bisect_cat.json", "c2_qft5_3qpu_cat.json", "c3_qft26_4qpu_cat.json.gz"):
        ops, meta = opstream.load(os.path.join(golden_dir, name))
        doc = opstream.encode(ops, meta)
        ops2, meta2 = opstream.decode(doc)
        assert meta == meta2
        assert opstream.encode(ops2, meta2) == doc


def test_c1_fixture_matches_reference_compiler_golden(golden_dir):
    """Shape pinned by the reference's tests/unit_tests/software/test_compiler.py:54-75 and the
    SURVEY Appendix D probe of the bisected GHZ-5."""
    ops, _ = opstream.load(os.path.join(golden_dir, "c1_ghz5_2qpu_bisect_cat.json"))
    n0 = [tuple(getattr(e, "name", e) if not isinstance(e, list) else tuple(e) for e in op) for op in ops["node_0"][0]]
    assert n0 == [("init", (0, 1, 2, 3, 4)), ("node_1", "cat", "correct"), ("cnot_gate", -1, 4),
                  (4, "node_1", "cat", "disentangle_start"), ("cnot_gate", 4, 3), ("cnot_gate", 3, 2)]
    n1 = [tuple(getattr(e, "name", e) if not isinstance(e, list) else tuple(e) for e in op) for op in ops["node_1"][0]]
    assert n1 == [("init", (0, 1, 2, 3)), ("h_gate", 3), ("cnot_gate", 3, 2), (2, "node_0", "cat", "entangle"),
                  (2, "node_0", "cat", "disentangle_end")]


_REF_EXAMPLES = "/root/reference/examples"


@pytest.mark.skipif(not os.path.isdir(_REF_EXAMPLES), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("kind", ["ghz", "qft"])
def test_synthetic_circuit_equals_reference_qasm_at_5_qubits(kind):
    def body(text):
        return [ln.strip() for ln in text.splitlines()
                if ln.strip() and not ln.startswith(("//", "OPENQASM", "include", "qreg", "creg"))]

    ours = (circuits.ghz_qasm if kind == "ghz" else circuits.qft_qasm)(5)
    with open(f"{_REF_EXAMPLES}/{kind}_5qubits.qasm") as f:
        ref = f.read()
    assert body(ours) == body(ref)


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "dqc_engine.h")).read()
    declared = set(re.findall(r"\b(dqe_[a-z0-9_]+)\s*\(", hdr))
    declared.discard("dqe_t")
    assert declared == set(engine.ABI), declared ^ set(engine.ABI)
    lib = engine.load_library()
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.dqe_version()


def test_create_reports_error_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(engine.EngineError):
        engine.Engine("ket", 3, 1)


# ---- gate-queue flusher (pass planner + peephole) on plan-only handles: host logic, no GPU ------
def _plan(circuit, **opts):
    from dqc_simulator_b200 import executor, experiment

    ops, _ = experiment.load_c2(circuit)
    e = engine.Engine("dm", 7, 1, device=-1)
    for k, v in opts.items():
        e.set_option(k, v)
    ex = executor.DQCExecutor(ops, experiment.c2_hardware(), e).run()
    ex.axes_of(ex.processing_qubits())
    e.flush()
    return e.stats(), ex


def test_planner_measurement_bound_pass_count():
    """Every measurement needs a per-shot global sum, so it ends a pass; everything between two
    measurements of the config-2 circuits fuses into one pass (tile = 12 of the 14 dm bits)."""
    st, ex = _plan("grover", tile_bits=12, cluster=0)
    assert ex.measurements == 288 and st["finish_launches"] == 288
    assert 288 <= st["passes"] <= 288 + 20
    assert st["api_ops"] > 3000 and st["merged_ops"] > 500
    st_q, ex_q = _plan("qft", tile_bits=12, cluster=0)
    assert ex_q.measurements == 44 and 44 <= st_q["passes"] <= 50
    st11, _ = _plan("grover", tile_bits=11, cluster=0)     # a smaller tile holds fewer high bits: more cuts
    assert st["passes"] < st11["passes"] < 2 * st["passes"]
    # cluster mode (default): the <= 8 tiles of a shot form a thread-block cluster, the measurement
    # is a micro-op (DSMEM sum + on-chip draw) and no longer cuts the pass
    stc, exc = _plan("grover", tile_bits=11)
    assert exc.measurements == 288 and stc["finish_launches"] == 0 and stc["passes"] < 0.6 * st11["passes"]
    std, _ = _plan("grover")        # default for a 14-bit register: 2^10 tiles, 16-CTA clusters
    assert std["finish_launches"] == 0 and std["passes"] < st11["passes"]


def test_planner_switches():
    full, _ = _plan("qft", cluster=0)
    nofuse, _ = _plan("qft", fuse=0)
    nomerge, _ = _plan("qft", merge=0, group=0, cluster=0)
    assert nofuse["passes"] > 5 * full["passes"]          # one op per pass
    assert nomerge["merged_ops"] == 0 and full["merged_ops"] > 100
    assert abs(nomerge["passes"] - full["passes"]) <= 4     # merging changes sweeps, hardly the pass cuts
    # amplitude-updates per pass are bounded by the live register (4^7 per shot)
    assert full["amp_updates"] <= full["passes"] * 4 ** 7


def test_plan_only_handle_refuses_reads():
    e = engine.Engine("ket", 4, 2, device=-1)
    e.alloc_axis(0)
    e.apply_1q(0, np.array([[0, 1], [1, 0]]))
    e.flush()
    with pytest.raises(engine.EngineError, match="plan-only"):
        e.full_state()


def test_oracle_reproduces_reference_printed_fidelities():
    """examples/run_experiment.py:89-93 prints (never asserts) end-to-end noisy fidelities from real NetSquid
    runs.  GHZ-5 is pinned digit for digit by trajectory enumeration (test_reference_printed_fidelities.py);
    QFT-5 and Grover-5 have 44 / 288 mid-circuit measurements, so here (a) trajectories that follow their own
    outcomes like the reference (host_branching, batch 1) must contain the printed value inside their
    spread, and (b) the batched path -- one shared timeline, predicated corrections advancing it by their
    expected duration -- must give a batch mean within 5e-5 of it (it was 2.4e-4 / 1.7e-4 low while every
    predicated correction was charged its full 135 us)."""
    from dqc_simulator_b200 import executor, experiment
    from oracle.numpy_engine import OracleEngine

    printed = {"ghz": 0.9570522876374276, "qft": 0.6470482939691747, "grover": 0.1071574284590638}
    mk = lambda f, n, b: OracleEngine(f, n, b, seed=0)
    for circuit, ntraj, nbatch in (("ghz", 8, 16), ("qft", 16, 32), ("grover", 4, 8)):
        want = printed[circuit]
        ops, _ = experiment.load_c2(circuit)
        data, ideal = experiment.ideal_output_ket(ops, mk)
        fs = []
        for s in range(ntraj):
            be = OracleEngine("dm", 7, 1, seed=100 + s)
            ex = executor.DQCExecutor(ops, experiment.c2_hardware(noisy=True), be, host_branching=True).run()
            fs.append(be.fidelity_ket(ex.axes_of(data), ideal, 0))
        fs = np.array(fs)
        assert abs(want - fs.mean()) < 4 * fs.std() + 1e-6 and abs(want - fs.mean()) < 1.5e-4, (circuit, fs.mean(), fs.std())
        be = OracleEngine("dm", 7, nbatch, seed=5)
        ex = executor.DQCExecutor(ops, experiment.c2_hardware(noisy=True), be).run()
        fb = np.mean([be.fidelity_ket(ex.axes_of(data), ideal, s) for s in range(nbatch)])
        assert abs(fb - want) < 5e-5, (circuit, fb, want)


def test_host_philox_matches_oracle_stream():
    from dqc_simulator_b200 import philox_host

    for seed, shot, draw in [(0, 0, 0), (2024, 7, 3), (2 ** 63 + 5, 2 ** 33 + 1, 2 ** 32 + 9)]:
        u, u2 = philox.shot_uniforms(seed, np.array([shot]), draw)
        hu, hu2 = philox_host.shot_uniforms(seed, shot, draw)
        assert hu == float(u[0]) and hu2 == float(u2[0])


def test_fidelity_squared_ket_and_uhlmann():
    from dqc_simulator_b200.experiment import fidelity_squared

    rng = np.random.default_rng(0)
    v = rng.normal(size=4) + 1j * rng.normal(size=4)
    v /= np.linalg.norm(v)
    a = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
    rho = a @ a.conj().T
    rho /= np.trace(rho)
    pure = np.outer(v, v.conj())
    assert abs(fidelity_squared(rho, v) - fidelity_squared(rho, pure)) < 1e-7      # Uhlmann reduces to <v|rho|v>
    assert abs(fidelity_squared(rho, rho) - 1) < 1e-10
    b = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
    sig = b @ b.conj().T
    sig /= np.trace(sig)
    assert abs(fidelity_squared(rho, sig) - fidelity_squared(sig, rho)) < 1e-10      # symmetric
    assert 0 < fidelity_squared(rho, sig) < 1


def _dm_plan_lines(capfd, monkeypatch, opts, tail_h=False):
    """Plan-only handle (no GPU): queue memory noise -> H -> gate noise -> CNOT -> 2-qubit depolarising on a
    qubit pair and return the planner's pass dump (DQE_DUMP_PLAN, written by the C library on fd 2)."""
    monkeypatch.setenv("DQE_DUMP_PLAN", "1")
    h = np.array([[1, 1], [1, -1]]) / np.sqrt(2)
    cnot = np.eye(4)[[0, 1, 3, 2]]
    e = engine.Engine("dm", 4, 1, device=-1)
    for k, v in opts.items():
        e.set_option(k, v)
    for a in (0, 1):
        e.alloc_axis(a)
    e.depolarize([0], 0.01); e.apply_1q(0, h); e.depolarize([0], 0.02)
    e.apply_2q(0, 1, cnot); e.depolarize([0, 1], 0.03)
    if tail_h:
        e.apply_1q(1, h)
    capfd.readouterr()
    e.flush()
    st = e.stats()
    return [ln.strip() for ln in capfd.readouterr().err.splitlines() if ln.startswith("pass ")], st


def test_planner_groups_factor_and_hoist_cx(capfd, monkeypatch):
    """Host logic of the gate-queue flusher, pinned on the plan it emits: the peephole folds noise -> H ->
    noise into one op (2 merges) but keeps its factors, so the op runs inside the qubit-pair group as
    Pauli / U rho U^dagger / Pauli; the CNOT commutes behind the 2-qubit depolarising channel and rides in
    the store addresses of the sweep ([>x]); one pass, one group."""
    lines, st = _dm_plan_lines(capfd, monkeypatch, {})
    assert lines == ["pass t=4 outer=0 nops=5: [>x] GROUP(2,3,0,1) gPAULI1.0 gMAT1X2.0 gPAULI1.0 gDEPOL2.0"]
    assert st["passes"] == 1 and st["merged_ops"] == 2 and st["micro_ops"] == 5


def test_planner_keeps_dense_blocks_out_of_groups(capfd, monkeypatch):
    """factor = 0: the composed 4x4 superoperator stays a stand-alone sweep (never a group sub-op)."""
    lines, _ = _dm_plan_lines(capfd, monkeypatch, {"factor": 0})
    assert lines == ["pass t=4 outer=0 nops=3: MAT2(2,0) [>x] GROUP(3,2,1,0) gDEPOL2.0"]


def test_planner_cx_in_the_middle_of_a_group(capfd, monkeypatch):
    """cxperm = 0 keeps queue order (CX before DEPOL2, as a sub-op); with cxperm a CX followed by an op on
    its control axis cannot move to the tail and stays a (thread-private round trip) sub-op."""
    lines, _ = _dm_plan_lines(capfd, monkeypatch, {"cxperm": 0})
    assert lines == ["pass t=4 outer=0 nops=6: GROUP(2,3,0,1) gPAULI1.0 gMAT1X2.0 gPAULI1.0 gMAT1X2.1 gDEPOL2.0"]
    lines, _ = _dm_plan_lines(capfd, monkeypatch, {}, tail_h=True)
    assert lines == ["pass t=4 outer=0 nops=7: GROUP(2,3,0,1) gPAULI1.0 gMAT1X2.0 gPAULI1.0 gDEPOL2.0 gMAT1X2.1 gMAT1X2.1"]


def test_data_collector_frame_matches_reference_shape():
    """util/helper.py:112-163: dc.dataframe["fidelity"].iloc[0] -- the read-out of every reference example
    (examples/getting_started.py:121-125).  Run with the trajectory that produced the printed number."""
    import ref_cases as rc
    from dqc_simulator_b200 import executor, experiment, hardware
    from oracle.numpy_engine import OracleEngine

    ops, _ = opstream.load(os.path.join(os.path.dirname(__file__), "golden", "ref_getting_started.json"))
    dqc = hardware.DQCSpec.uniform(3, state4distribution=hardware.werner_state(0.9), num_positions=10,
                                   num_comm_qubits=2, **rc.C2_NOISE)
    bell = np.array([1, 0, 0, 1]) / np.sqrt(2)
    frames = []
    for ref in (bell.reshape(4, 1), np.outer(bell, bell)):   # ket column vector (as in the example) or dm
        be = OracleEngine("dm", 7, 1, seed=1)
        ex = executor.DQCExecutor(ops, dqc, be, forced_outcomes=[1, 0] + [0] * 8, host_branching=True).run()
        dc = experiment.get_data_collector(ex, [(2, "node_0"), ([2], "node_1")], ref)
        frames.append(dc.dataframe)
    for df in frames:
        assert list(df.columns) == ["time_stamp", "entity_name", "fidelity"] and len(df) == 1
        assert df["time_stamp"].iloc[0] > 0 and df["entity_name"].iloc[0] == "DQCMasterProtocol"
    assert abs(frames[0]["fidelity"].iloc[0] - 0.8921630426886507) < 2e-15
    assert abs(frames[1]["fidelity"].iloc[0] - 0.8921630426886507) < 1e-7   # Uhlmann through sqrtm-free eigh
    with pytest.raises(TypeError, match="not of correct form"):
        experiment.get_data_collector(ex, [("2", "node_0")], bell)


def test_program_instructions_overlap_on_free_positions():
    """QPU.execute_program schedules like a NetSquid ``QuantumProgram(parallel=True)`` on physical instructions
    with ``parallel=True`` (quantum_processors.py:344-456): in-order issue, an instruction starts when its
    positions are free; INIT is exclusive.  Pinned on the simulated clock (ns)."""
    from dqc_simulator_b200 import executor, hardware, instructions as ins
    from oracle.numpy_engine import OracleEngine

    def run(parallel):
        dqc = hardware.DQCSpec.uniform(1, num_positions=6, num_comm_qubits=2)
        ex = executor.DQCExecutor({"node_0": [[]]}, dqc, OracleEngine("ket", 6, 1), parallel_programs=parallel)
        q = ex.qpus["node_0"]
        prog = executor.QuantumProgram()
        prog.apply(ins.INSTR_INIT, [2, 3, 4])          # 3 ns, exclusive
        prog.apply(ins.INSTR_H, [2])                   # 135 us
        prog.apply(ins.INSTR_H, [3])                   # overlaps with the first H
        prog.apply(ins.INSTR_CNOT, [2, 3])             # 600 us, waits for both
        prog.apply(ins.INSTR_X, [4])                   # free position, but issued after the CNOT has started
        q.execute_program(prog)
        return q.clock

    assert run(False) == 3 + 135e3 + 135e3 + 600e3 + 135e3
    assert run(True) == 3 + 135e3 + 600e3              # X [135e3, 270e3] runs under the CNOT [135e3, 735e3]


@pytest.mark.parametrize("case,n,want", [
    # per-shot timelines (round 2): the clock arithmetic rides along as classical instructions; cat-entangle runs
    # (inject -> CNOT -> measure-and-discard) and measure-out runs (CNOT -> H -> measure-and-discard) are fused at flush,
    # and whole remote gates become two ops on the data pair: NO comm qubit ever goes live, every pass runs on the
    # 5-qubit rho (Grover-5 was 292 passes / 4209 micro-ops / 3.0 M element updates per shot before these rewrites,
    # 0.15 M now)
    ("c2_grover5_3qpu_cat.json", 7, dict(passes=144, micro_ops=1832, api_ops=5313, merged_ops=2229)),
    ("c2_qft5_3qpu_cat.json", 7, dict(passes=22, micro_ops=271, api_ops=812, merged_ops=355)),
    ("c4_qft16_mono.json.gz", 16, dict(passes=56, micro_ops=1182, api_ops=1564, merged_ops=521)),
])
def test_planner_pass_counts_of_the_benchmark_circuits(case, n, want):
    """Plan-only handles (no GPU): the pass / micro-op counts the measured numbers in profiles/ were taken
    with (round 2: every shot follows its own simulated timeline; the monolithic C4 circuit has no
    measurement, so all of its times stay host constants).  A planner change that alters them should be a
    conscious one."""
    from dqc_simulator_b200 import executor, experiment, hardware

    ops, _ = opstream.load(os.path.join(os.path.dirname(__file__), "golden", case))
    if n == 16:
        dqc = hardware.DQCSpec.uniform(1, num_positions=16, num_comm_qubits=0, **experiment.C2_NOISE)
    else:
        dqc = experiment.c2_hardware(noisy=True)
    e = engine.Engine("dm", n, 1000 if n == 7 else 1, device=-1)
    executor.DQCExecutor(ops, dqc, e).run()
    e.flush()
    st = e.stats()
    assert {k: st[k] for k in want} == want
    assert st["finish_launches"] == 0      # every measurement of these circuits is fused into its pass


# ---- per-shot simulated timelines -------------------------------------------------
def _per_shot_case(circuit, batch, seed, make=None, check=None):
    """A sampled batch in which every shot follows its own timeline, and the same shots re-run one at a
    time with the outcomes forced (the host then knows every outcome and branches exactly like the
    reference does: dqc_control.py:347-348, 385-403, 497-503; physical_layer.py:269-335)."""
    from dqc_simulator_b200 import executor, experiment
    from oracle.numpy_engine import OracleEngine

    ops, _ = experiment.load_c2(circuit)
    dqc = experiment.c2_hardware(noisy=True)
    be = (make or OracleEngine)("dm", 7, batch, seed=seed)
    if hasattr(be, "set_option"):
        be.set_option("outcome_log", 512)
    ex = executor.DQCExecutor(ops, dqc, be).run()
    axes = ex.axes_of(ex.processing_qubits())
    rho = [be.reduced_dm(axes, s) for s in range(batch)]
    outcomes = be.read_outcomes()
    end = ex.read_time(ex.end_time)
    singles = {}
    for s in range(batch if check is None else check):
        b1 = OracleEngine("dm", 7, 1, seed=seed, shot_offset=s)
        e1 = executor.DQCExecutor(ops, dqc, b1, forced_outcomes=list(outcomes[:, s])).run()
        singles[s] = (b1.reduced_dm(e1.axes_of(e1.processing_qubits()), 0), e1.end_time)
    return rho, end, singles, outcomes, ex


@pytest.mark.parametrize("circuit,batch", [("ghz", 16), ("qft", 12)])
def test_batched_shots_follow_their_own_timeline(circuit, batch):
    """Shot by shot, a sampled batch equals the single-trajectory run with the same outcomes: rho to 1e-12
    and the simulated end time exactly -- the batch no longer shares one (expected) timeline."""
    rho, end, singles, outcomes, ex = _per_shot_case(circuit, batch, seed=11)
    assert len({tuple(outcomes[:, s]) for s in range(batch)}) > 1          # the shots really differ
    for s in range(batch):
        assert np.abs(rho[s] - singles[s][0]).max() < 1e-12
        assert abs(end[s] - singles[s][1]) < 1e-5      # ns: (register + host offset) vs plain float rounding
    assert len(set(end.tolist())) > 1                                      # ... and so do their timelines
    assert ex.T.peak <= 128


def test_per_shot_timeline_reproduces_the_trajectory_spread():
    """The per-shot fidelity spread of a QFT-5 batch is the trajectory spread (6.8e-5 from single-trajectory
    runs), not the 2.5e-5 of the shared expected timeline of round 1."""
    from dqc_simulator_b200 import executor, experiment
    from oracle.numpy_engine import OracleEngine

    ops, _ = experiment.load_c2("qft")
    data, ideal = experiment.ideal_output_ket(ops, lambda f, n, b: OracleEngine(f, n, b))
    spreads = {}
    for exact in (True, False):
        be = OracleEngine("dm", 7, 64, seed=3)
        ex = executor.DQCExecutor(ops, experiment.c2_hardware(noisy=True), be, per_shot_time=exact).run()
        axes = ex.axes_of(data)
        f = np.array([be.fidelity_ket(axes, ideal, s) for s in range(64)])
        spreads[exact] = f.std()
        assert abs(f.mean() - 0.6470482939691747) < 4 * 6.8e-5 / 8 + 1e-5
    assert 4.5e-5 < spreads[True] < 9.5e-5 and spreads[False] < 0.6 * spreads[True]


def test_reinit_advances_the_shot_ids():
    """dqe_reinit must not replay the random stream (plan-only handle: host logic)."""
    e = engine.Engine("dm", 3, 4, seed=5, device=-1)
    e.set_option("shot_stride", 2)
    e.reinit()     # offset 0 -> 8
    e.reinit()     # -> 16: no error, the arithmetic itself is covered by the GPU twin


# ---- fused remote-gate measurements: the host algebra, checked without a GPU ---------------------
def _fused_tables(capfd, monkeypatch, build):
    """Queue a program on a plan-only handle and return the fused measurement ops the flusher wrote
    (DQE_DUMP_QUEUE: 'Q type=6 ... nb=4 ... b=rowA,rowY,colA,colY,A,Y,form,real c=<constants>')."""
    monkeypatch.setenv("DQE_DUMP_QUEUE", "1")
    e = engine.Engine("dm", 4, 1, device=-1, tile_bits=6)
    build(e)
    capfd.readouterr()
    e.flush()
    out = []
    for ln in capfd.readouterr().err.splitlines():
        if ln.startswith("Q type=6") and " nb=4 " in ln:
            vals = [complex(*map(float, x.split(","))) for x in ln.split(" c=")[1].strip().split(";")]
            out.append(np.array(vals))
    return out, e.stats()


def _unpack(vals, nt, rows, cols):
    """(t_0, t_1, T_0, T_1) from the staged constants: complex tables, or real ones packed two doubles to an entry."""
    t = vals[:2 * nt].reshape(2, nt)
    body = vals[2 * nt:]
    if len(body) == rows * cols:          # real-packed: 2 outcomes x rows x cols doubles
        flat = np.stack([body.real, body.imag], axis=1).reshape(-1)
        return t, flat.reshape(2, rows, cols).astype(complex)
    return t, body.reshape(2, rows, cols)


@pytest.mark.parametrize("complex_gate", [False, True])
def test_fused_injecting_run_tables_equal_a_direct_simulation(capfd, monkeypatch, complex_gate):
    """fuse_cat_measure (dqe_engine.cu): INJECT(x, y) -> [U on x] -> CNOT(A -> x) -> 2-qubit depolarising -> measure x
    (flip e, discard) becomes ONE op with tables T_m (16 x 4) and functionals t_m.  Every column of T_m must equal the
    direct density-matrix simulation of the same run on the matrix unit |rA'><cA'| of A (x) the injected pair."""
    from dqc_simulator_b200 import hardware

    W = np.asarray(hardware.werner_state(0.93), dtype=complex).reshape(4, 4)
    p2, flip = 0.07, 0.04
    cnot = np.eye(4)[[0, 1, 3, 2]]
    rng = np.random.default_rng(3)
    z = rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2))
    U = np.linalg.qr(z)[0] if complex_gate else np.eye(2)
    A, x, y = 0, 3, 2

    def build(e):
        e.alloc_axis(A)
        e.apply_1q(A, np.array([[1, 1], [1, -1]]) / np.sqrt(2))
        e.inject_2q(x, y, W, True)
        if complex_gate:
            e.apply_1q(x, U)
        e.apply_2q(A, x, cnot)
        e.depolarize([A, x], p2)
        e.measure(x, 0, -1, True, flip)

    ops, st = _fused_tables(capfd, monkeypatch, build)
    assert len(ops) == 1 and st["passes"] == 1
    t, T = _unpack(ops[0], 4, 16, 4)
    assert len(ops[0]) == (8 + 128 if complex_gate else 8 + 64)   # real tables are staged as packed doubles
    # direct simulation on (A, x, y), A = most significant
    I2 = np.eye(2)
    CX = np.kron(cnot, I2)                       # control A, target x
    Ux = np.kron(np.kron(I2, U), I2)
    for j in range(4):
        rA, cA = j >> 1, j & 1
        E = np.zeros((2, 2), complex); E[rA, cA] = 1.0
        rho = np.kron(E, W)                      # W indexed (x y)
        rho = Ux @ rho @ Ux.conj().T
        rho = CX @ rho @ CX.conj().T
        r6 = rho.reshape(2, 2, 2, 2, 2, 2)       # rA rx ry cA cx cy
        red = np.einsum("axyaxz->yz", r6)        # Tr over (A, x): what is left on y
        mixed = np.einsum("ab,xc,yz->axybcz", I2 / 2, I2 / 2, red)
        r6 = (1 - p2) * r6 + p2 * mixed
        for m in range(2):
            kept = (1 - flip) * r6[:, m, :, :, m, :] + flip * r6[:, 1 - m, :, :, 1 - m, :]   # rA ry cA cy
            assert np.abs(T[m][:, j] - kept.reshape(16)).max() < 1e-14
            assert abs(t[m][j] - np.einsum("ayay->", kept)) < 1e-14


def test_fused_measure_out_run_tables_equal_a_direct_simulation(capfd, monkeypatch):
    """Contracting form: constant ops on (A, x) right in front of a measure-and-discard of an EXISTING axis x fold into
    the measurement (4 x 16 tables); per-shot or predicated ops end the run and stay outside."""
    cnot = np.eye(4)[[0, 1, 3, 2]]
    h = np.array([[1, 1], [1, -1]]) / np.sqrt(2)
    p2, p1, flip = 0.05, 0.02, 0.01
    A, x = 1, 2

    def build(e):
        for a in (A, x):
            e.alloc_axis(a)
        e.apply_1q(A, h)
        e.cond_1q(x, np.array([[0, 1], [1, 0]]), 5)      # predicated: must stay outside the fused op
        e.apply_2q(x, A, cnot)                           # control x, target A
        e.depolarize([x, A], p2)
        e.apply_1q(x, h)
        e.depolarize([x], p1)
        e.measure(x, 1, -1, True, flip)

    ops, st = _fused_tables(capfd, monkeypatch, build)
    assert len(ops) == 1
    t, T = _unpack(ops[0], 16, 4, 16)
    CX = np.zeros((4, 4))
    for a in range(2):
        for b in range(2):
            CX[(a ^ b) * 2 + b, a * 2 + b] = 1.0         # index (A x): control x, target A
    Hx = np.kron(np.eye(2), h)
    for i in range(16):
        rA, rx, cA, cx = (i >> 3) & 1, (i >> 2) & 1, (i >> 1) & 1, i & 1
        rho = np.zeros((4, 4), complex); rho[rA * 2 + rx, cA * 2 + cx] = 1.0
        rho = CX @ rho @ CX.T
        rho = (1 - p2) * rho + p2 * np.trace(rho) * np.eye(4) / 4
        rho = Hx @ rho @ Hx.T
        r4 = rho.reshape(2, 2, 2, 2)                     # rA rx cA cx
        r4 = (1 - p1) * r4 + p1 * np.einsum("axbx->ab", r4)[:, None, :, None] * (np.eye(2) / 2)[None, :, None, :]
        for m in range(2):
            kept = (1 - flip) * r4[:, m, :, m] + flip * r4[:, 1 - m, :, 1 - m]
            assert np.abs(T[m][:, i] - kept.reshape(4)).max() < 1e-14
            assert abs(t[m][i] - np.trace(kept)) < 1e-14


@pytest.mark.parametrize("two_dyn", [False, True])
def test_whole_remote_gate_tables_equal_a_direct_simulation(capfd, monkeypatch, two_dyn):
    """catfuse = 3 (try_fuse_whole_at): a cat-comm remote gate -- pair (x, y), CNOT(A -> x), measure x; X on y if the
    outcome was 1, per-shot idle noise D(q) on y; CNOT(y -> B), H(y), measure y -- becomes two ops on the DATA pair
    (A, B); neither comm qubit goes live.  The 16 x 16 matrix C_m2 = T2_m2 (K0 + q K1) the kernel builds from the staged
    tables must equal the direct density-matrix simulation for every (m1, m2) and two strengths q."""
    from dqc_simulator_b200 import hardware, timeline as tl

    W = np.asarray(hardware.werner_state(0.95), dtype=complex).reshape(4, 4)
    cnot = np.eye(4)[[0, 1, 3, 2]]
    Hm = np.array([[1, 1], [1, -1]]) / np.sqrt(2)
    X = np.array([[0, 1], [1, 0]], dtype=float)
    p2a, p2b, p1, f1, f2, px = 0.03, 0.02, 0.01, 0.004, 0.006, 0.015
    A, B, x, y = 0, 1, 3, 2
    monkeypatch.setenv("DQE_DUMP_QUEUE", "1")
    e = engine.Engine("dm", 4, 1, device=-1, tile_bits=6)
    for a in (A, B):
        e.alloc_axis(a)
        e.apply_1q(a, Hm)
    e.apply_2q(A, B, cnot)
    e.classical_op(tl.CL_CONST, 5, imm=0.25)
    e.inject_2q(x, y, W, True)
    e.apply_2q(A, x, cnot); e.depolarize([A, x], p2a)
    e.measure(x, 0, -1, True, f1)
    e.set_predicate(0); e.depolarize([y], px); e.apply_1q(y, X); e.set_predicate(-1)
    e.dyn_channel(y, tl.DYN_DEPOL, 5)
    if two_dyn:
        e.classical_op(tl.CL_CONST, 6, imm=0.125)
        e.dyn_channel(y, tl.DYN_DEPHASE, 6)       # a second per-shot channel: the tables become bilinear in (q1, q2)
    e.apply_2q(y, B, cnot); e.depolarize([y, B], p2b); e.apply_1q(y, Hm); e.depolarize([y], p1)
    e.measure(y, 1, -1, True, f2)
    capfd.readouterr()
    e.flush()
    err = capfd.readouterr().err
    fused = [ln for ln in err.splitlines() if ln.startswith("Q type=6") and " nb=4 " in ln]
    assert len(fused) == 2 and e.stats()["passes"] == 1
    wl = [ln for ln in fused if ln.split(" c=")[1].count(";") > 100][0]
    aux = int(wl.split(" aux=")[1].split()[0])
    assert aux & 0xff == 0 and ((aux >> 8) & 0xff) - 1 == 0 and ((aux >> 16) & 0xff) - 1 == 5   # m1 bit, predicate, register
    assert ((aux >> 24) & 0xff) - 1 == (6 if two_dyn else -1)
    vals = np.array([complex(*map(float, v.split(","))) for v in wl.split(" c=")[1].strip().split(";")])
    flat = np.stack([vals.real, vals.imag], axis=1).reshape(-1)
    nord = 4 if two_dyn else 2
    # the predicate IS the first outcome here: the table ships compact, K[m1][order] (= K[m1][pred = m1][order])
    Kc = flat[:128 * nord].reshape(2, nord, 16, 4)               # m1, order, (rA cA ry cy), (rA' cA')
    K = np.stack([Kc, Kc], axis=1)                               # indexed [m1, m1, order] below
    T2 = flat[128 * nord:128 * nord + 128].reshape(2, 4, 16)     # m2, (rB cB), (rB' ry cB' cy)
    q2 = 0.21 if two_dyn else 0.0

    def chan2(rho, p, keep_shape):                  # two-qubit depolarising on the first two (of 3 or 4) qubits
        d = rho.shape[0] // 4
        r = rho.reshape(4, d, 4, d)
        red = np.einsum("akal->kl", r)
        return ((1 - p) * r + p * np.einsum("ab,kl->akbl", np.eye(4) / 4, red)).reshape(rho.shape)

    def kron(*ms):
        out = np.eye(1)
        for m in ms:
            out = np.kron(out, m)
        return out

    I2 = np.eye(2)
    for m1 in range(2):
        for m2 in range(2):
            for q in (0.0, 0.37):
                C = np.zeros((16, 16))
                for eo in range(16):
                    rA, rB, cA, cB = (eo >> 3) & 1, (eo >> 2) & 1, (eo >> 1) & 1, eo & 1
                    for ei in range(16):
                        rA2, rB2, cA2, cB2 = (ei >> 3) & 1, (ei >> 2) & 1, (ei >> 1) & 1, ei & 1
                        acc = 0.0
                        for yy in range(4):
                            Ke = K[m1, m1, 0] + q * K[m1, m1, 1]
                            if two_dyn:
                                Ke = Ke + q2 * (K[m1, m1, 2] + q * K[m1, m1, 3])
                            k = Ke[(rA * 2 + cA) * 4 + yy, rA2 * 2 + cA2]
                            acc += T2[m2, rB * 2 + cB, rB2 * 8 + (yy >> 1) * 4 + cB2 * 2 + (yy & 1)] * k
                        C[eo, ei] = acc
                for ei in range(16):                 # matrix unit of (A, B) -> the direct simulation
                    rA2, rB2, cA2, cB2 = (ei >> 3) & 1, (ei >> 2) & 1, (ei >> 1) & 1, ei & 1
                    Eu = np.zeros((4, 4)); Eu[rA2 * 2 + rB2, cA2 * 2 + cB2] = 1.0
                    # qubit order (A, x, y, B): start from (A, B) (x) W(x, y) and move B last
                    rho = np.einsum("abcd,xyzw->axybczwd", Eu.reshape(2, 2, 2, 2), W.reshape(2, 2, 2, 2).real).reshape(16, 16)
                    U = kron(cnot, I2, I2); rho = U @ rho @ U.T          # CNOT(A -> x)
                    rho = chan2(rho, p2a, None)                           # depolarising on (A, x)
                    r = rho.reshape(2, 2, 4, 2, 2, 4)                     # A x (yB) | A x (yB)
                    rho = ((1 - f1) * r[:, m1, :, :, m1, :] + f1 * r[:, 1 - m1, :, :, 1 - m1, :]).reshape(8, 8)   # (A, y, B)
                    if m1:                                                # predicated: noise then X on y
                        r = rho.reshape(2, 2, 2, 2, 2, 2)
                        red = np.einsum("aybcyd->abcd", r)
                        r = (1 - px) * r + px * np.einsum("abcd,yz->aybczd", red, I2 / 2)
                        rho = r.reshape(8, 8)
                        U = kron(I2, X, I2); rho = U @ rho @ U.T
                    r = rho.reshape(2, 2, 2, 2, 2, 2)                     # D(q) on y
                    red = np.einsum("aybcyd->abcd", r)
                    r = (1 - q) * r + q * np.einsum("abcd,yz->aybczd", red, I2 / 2)
                    if two_dyn:                                           # dephasing of strength q2 on y
                        r = r.copy()
                        r[:, 0, :, :, 1, :] *= 1 - q2
                        r[:, 1, :, :, 0, :] *= 1 - q2
                    rho = r.reshape(8, 8)
                    # CNOT(y -> B), depolarising on (y, B), H on y, depolarising on y; order (A, y, B)
                    U = kron(I2, cnot); rho = U @ rho @ U.T
                    r = rho.reshape(2, 4, 2, 4)
                    red = np.einsum("akbk->ab", r)
                    rho = ((1 - p2b) * r + p2b * np.einsum("ab,kl->akbl", red, np.eye(4) / 4)).reshape(8, 8)
                    U = kron(I2, Hm, I2); rho = U @ rho @ U.T
                    r = rho.reshape(2, 2, 2, 2, 2, 2)
                    red = np.einsum("aybcyd->abcd", r)
                    r = (1 - p1) * r + p1 * np.einsum("abcd,yz->aybczd", red, I2 / 2)
                    kept = (1 - f2) * r[:, m2, :, :, m2, :] + f2 * r[:, 1 - m2, :, :, 1 - m2, :]    # rA rB cA cB
                    assert np.abs(C[:, ei] - kept.reshape(16)).max() < 1e-14, (m1, m2, q, ei)


def test_recorded_call_stream_replays_the_same_plan():
    """Engine.record / play (host-side plan cache for programs that span several flushes, e.g. a sharded state
    vector's passes - exchange - passes): the replayed C-ABI stream plans exactly what the walked program planned, and
    a run that reads device data back, or changes an option on the way, refuses to be replayed."""
    from dqc_simulator_b200 import executor, hardware

    ops, _ = opstream.load(os.path.join(os.path.dirname(__file__), "golden", "c5_qft12_mono.json"))
    dqc = hardware.DQCSpec.uniform(1, num_positions=12, num_comm_qubits=0)
    e = engine.Engine("ket", 12, 1, device=-1)
    with e.record() as rec:
        executor.DQCExecutor(ops, dqc, e).run()
        e.flush()
    assert rec.valid and len(rec) > 100
    want, live = e.stats(), e.live
    e.reinit()
    e.reset_stats()
    e.play(rec)
    got = e.stats()
    assert got == want and e.live == live
    with e.record() as bad:
        e.alloc_axis(0) if 0 not in e.live else None
        e.set_option("merge", 1)
    assert not bad.valid and bad.why == "dqe_set_option"
    with pytest.raises(engine.EngineError, match="cannot be replayed"):
        e.play(bad)
    other = engine.Engine("ket", 12, 1, device=-1)
    with pytest.raises(engine.EngineError, match="handle it was recorded on"):
        other.play(rec)


def test_netsquid_formalism_names_resolve_to_the_two_engines():
    """Every QFormalism member a NetSquid user can select lands on one of the two dense engines: SPARSEDM on the
    density-matrix one (same rho), STAB / GSLC on the state-vector one (every stabiliser state is a ket)."""
    for name, want in (("ket", engine.KET), ("KET", engine.KET), ("dm", engine.DM), ("SPARSEDM", engine.DM),
                       ("stab", engine.KET), ("GSLC", engine.KET), (0, engine.KET), (1, engine.DM)):
        assert engine.resolve_formalism(name) == want
    e = engine.Engine("sparsedm", 3, 1, device=-1)
    assert e.formalism == engine.DM
    e.close()
    with pytest.raises(ValueError, match="unknown state formalism"):
        engine.Engine("mps", 3, 1, device=-1)
