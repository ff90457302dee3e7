"""The reference's known-answer cases (tests/ref_cases.py) and the committed reference-compiled
circuits, run on the CUDA engine through the C-ABI, checked against the closed forms and against
the oracle on the same seeded inputs (1e-10 max-abs, classical bits bit-exact)."""
import glob
import os

import numpy as np
import pytest

import ref_cases as rc
from dqc_simulator_b200 import hardware
from oracle.numpy_engine import OracleEngine

pytestmark = pytest.mark.gpu
TOL = 1e-10


def gpu(formalism, max_qubits, batch):
    from dqc_simulator_b200.engine import Engine

    return Engine(formalism, max_qubits, batch, seed=1)


def cpu(formalism, max_qubits, batch):
    return OracleEngine(formalism, max_qubits, batch, seed=1)


def _cmp(got, want, tol=TOL):
    if isinstance(got, list):
        for g, w in zip(got, want):
            _cmp(g, w, tol)
        return
    assert np.abs(np.asarray(got) - np.asarray(want)).max() < tol


@pytest.mark.parametrize("case,args", [
    (rc.noise_entangling_gate_10_percent, ()), (rc.noise_entangling_gate_extra_qubit, (False,)),
    (rc.noise_entangling_gate_extra_qubit, (True,)), (rc.noise_recover_ideal_when_error_0, ()),
    (rc.noise_02_entangling_gate, (False,)), (rc.noise_02_entangling_gate, (True,)),
    (rc.memory_depol, (1, 0.0)), (rc.memory_depol, (2, 0.0)), (rc.memory_depol, (10, 0.0)),
    (rc.memory_depol, (1, 1e9)), (rc.qpu_mem_depol_by_position_type, ("comm",)),
    (rc.qpu_mem_depol_by_position_type, ("proc",)), (rc.qpu_cnot_depol_in_7_qubit_register, (0, 1)),
    (rc.qpu_cnot_depol_in_7_qubit_register, (2, 3))])
def test_reference_hardware_known_answers(case, args):
    _cmp(*case(gpu, *args))


@pytest.mark.parametrize("formalism", ["dm", "ket"])
@pytest.mark.parametrize("case", sorted(rc.CIRCUIT_CASES))
def test_reference_circuit_cases(case, formalism):
    name, groups, wants, _ = rc.CIRCUIT_CASES[case]
    ex, be, _ = rc.run_golden(gpu, name, formalism)
    for qubits, want in zip(groups, wants):
        if want is not None:
            _cmp(rc.rdm(ex, be, qubits), want)


@pytest.mark.parametrize("kind", ["local", "remote_cat"])
@pytest.mark.parametrize("a,b,t", [(0, 0, 0), (0, 1, 0), (1, 0, 0), (1, 1, 0), (1, 1, 1)])
def test_toffoli_truth_table(kind, a, b, t):
    ex, be, _ = rc.run_golden(gpu, f"ref_toffoli_{kind}_{a}{b}{t}.json")
    tq = (4, "node_0") if kind == "local" else (2, "node_1")
    _cmp(rc.rdm(ex, be, [(2, "node_0"), (3, "node_0"), tq]), rc.toffoli_want(a, b, t))


def test_logged_measurements_host_branching():
    ex, be, _ = rc.run_golden(gpu, "ref_logged_meas.json", host_branching=True)
    assert [r for _, _, r in ex.logged_results] == [1, 0, 1]


@pytest.mark.parametrize("forced", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("formalism", ["ket", "dm"])
def test_config1_ghz5_all_forced_outcomes(formalism, forced):
    ex, be, meta = rc.run_golden(gpu, "c1_ghz5_2qpu_bisect_cat.json", formalism, max_qubits=7, forced=forced,
                                 num_positions=5, num_comm_qubits=2)
    data = [tuple(meta["qubit_map"][str(i)]) for i in range(5)]
    ghz = np.zeros(32, dtype=complex)
    ghz[0] = ghz[31] = 2 ** -0.5
    _cmp(rc.rdm(ex, be, data), rc.ket2dm(ghz))


NOISE = dict(p_depolar_error_cnot=1e-3, single_qubit_gate_error_prob=2e-5, meas_error_prob=3e-3,
             comm_qubit_depolar_rate=0.055, proc_qubit_depolar_rate=0.055)


def _c2_dqc(extra=None):
    kw = dict(NOISE)
    kw.update(extra or {})
    return hardware.DQCSpec.uniform(3, state4distribution=hardware.werner_state(0.99), num_positions=10,
                                    num_comm_qubits=2, **kw)


@pytest.mark.parametrize("name", ["c2_ghz5_3qpu_cat.json", "c2_qft5_3qpu_cat.json", "c2_grover5_3qpu_cat.json"])
def test_config2_noisy_dm_engine_vs_oracle(name):
    """BASELINE config 2 at oracle size (4 shots): noisy DM, sampled outcomes from the shared Philox
    stream -> same trajectories, rho to 1e-10, creg bit-exact."""
    res = []
    for mk in (gpu, cpu):
        ex, be, meta = rc.run_golden(mk, name, "dm", max_qubits=7, batch=4, dqc=_c2_dqc())
        data = [tuple(v) for v in meta["qubit_map"].values()]
        axes = ex.axes_of(data)
        res.append((be.full_state(), be.read_creg_words() if mk is gpu else be.creg,
                    [be.reduced_dm(axes, s) for s in range(4)]))
    assert np.abs(res[0][0] - res[1][0]).max() < TOL
    assert np.array_equal(res[0][1], res[1][1])
    _cmp(res[0][2], res[1][2])


def test_config2_with_dephasing_and_amplitude_damping():
    """north_star channels not wired by the reference: oracle-only parity."""
    res = []
    for mk in (gpu, cpu):
        ex, be, _ = rc.run_golden(mk, "c2_qft5_3qpu_cat.json", "dm", max_qubits=7, batch=2,
                                  dqc=_c2_dqc(dict(mem_dephase_rate=0.2, mem_amp_damp_rate=0.1)))
        res.append(be.full_state())
    assert np.abs(res[0] - res[1]).max() < TOL


@pytest.mark.parametrize("name,nq,npos", [("c3_qft8_4qpu_cat.json", 10, 4), ("c3_ghz10_4qpu_cat.json", 12, 5),
                                          ("c3_qft12_4qpu_cat.json", 14, 5)])
def test_config3_shape_ket_engine_vs_oracle(name, nq, npos):
    """BASELINE config 3 shape at oracle size: QFT/GHZ over 4 QPUs, cat-comm, ket, ideal."""
    res = []
    for mk in (gpu, cpu):
        ex, be, meta = rc.run_golden(mk, name, "ket", max_qubits=nq, batch=2, num_qpus=4,
                                     num_positions=npos, num_comm_qubits=2)
        res.append((be.full_state(), be.read_creg_words() if mk is gpu else be.creg))
    assert np.abs(res[0][0] - res[1][0]).max() < TOL
    assert np.array_equal(res[0][1], res[1][1])


@pytest.mark.parametrize("name,formalism,nq", [("c4_qft6_mono.json", "dm", 6), ("c5_qft12_mono.json", "ket", 12),
                                                ("c5_ghz12_mono.json", "ket", 12)])
def test_config45_shape_engine_vs_oracle(name, formalism, nq):
    res = []
    noise = NOISE if formalism == "dm" else {}
    for mk in (gpu, cpu):
        ex, be, _ = rc.run_golden(mk, name, formalism, max_qubits=nq, batch=1, num_qpus=1,
                                  num_positions=nq, num_comm_qubits=0, **noise)
        res.append(be.full_state())
    assert np.abs(res[0] - res[1]).max() < TOL


def test_every_fixture_runs(golden_dir):
    """All small committed op streams execute on the engine without planner errors."""
    for path in sorted(glob.glob(os.path.join(golden_dir, "ref_*.json"))):
        ex, be, _ = rc.run_golden(gpu, os.path.basename(path))
        be.sync()
        assert ex.finished


@pytest.mark.parametrize("name", ["c2_qft5_3qpu_cat.json", "c2_ghz5_3qpu_cat.json"])
def test_config2_with_pinned_low_comm_axes(name):
    """Executor option comm_axes_low (comm qubits on the lowest, pinned axes: longer contiguous HBM
    runs per tile): same physics, different axis assignment -- engine vs oracle on the same option,
    and reduced states equal to the default assignment."""
    from dqc_simulator_b200 import executor, opstream

    ops, meta = opstream.load(os.path.join(rc.GOLDEN, name))
    data = [tuple(v) for v in meta["qubit_map"].values()]
    out = {}
    for label, mk, low in (("gpu_low", gpu, 2), ("cpu_low", cpu, 2), ("gpu_default", gpu, 0)):
        be = mk("dm", 7, 3)
        ex = executor.DQCExecutor(ops, _c2_dqc(), be, comm_axes_low=low).run()
        axes = ex.axes_of(data)
        out[label] = ([be.reduced_dm(axes, s) for s in range(3)], be.full_state())
    assert np.abs(out["gpu_low"][1] - out["cpu_low"][1]).max() < TOL
    _cmp(out["gpu_low"][0], out["cpu_low"][0])
    _cmp(out["gpu_low"][0], out["gpu_default"][0])


def test_data_collector_on_a_shot_batch():
    """experiment.get_data_collector (util/helper.py:112-163) on the CUDA engine: one frame row per shot, every
    row of a forced-outcome batch equal to the reference's printed getting_started fidelity."""
    from dqc_simulator_b200 import executor, experiment, opstream

    ops, _ = opstream.load(os.path.join(os.path.dirname(__file__), "golden", "ref_getting_started.json"))
    dqc = hardware.DQCSpec.uniform(3, state4distribution=hardware.werner_state(0.9), num_positions=10,
                                   num_comm_qubits=2, **rc.C2_NOISE)
    be = gpu("dm", 7, 5)
    # forced outcomes are known on the host: the executor branches like the reference (no predicated gates),
    # so every shot of the batch follows the trajectory that produced the printed number
    ex = executor.DQCExecutor(ops, dqc, be, forced_outcomes=[1, 0] + [0] * 8).run()
    df = experiment.get_data_collector(ex, [(2, "node_0"), (2, "node_1")], rc.B00).dataframe
    assert len(df) == 5 and list(df.columns) == ["time_stamp", "entity_name", "fidelity"]
    assert np.abs(df["fidelity"].to_numpy() - 0.8921630426886507).max() < 1e-13


@pytest.mark.parametrize("circuit,batch", [("ghz", 64), ("qft", 64), ("grover", 64)])
def test_batched_shots_follow_their_own_timeline_on_the_engine(circuit, batch):
    """a sampled batch (>= 64 shots) of the C2 circuits on the CUDA engine matches, shot by
    shot at 1e-10 in rho, single-trajectory runs forced to the same outcomes (the host then branches exactly like
    the reference: dqc_control.py:347-348, 385-403, 497-503; physical_layer.py:269-335), and the simulated end
    times agree: the batch no longer shares one expected timeline."""
    from test_host_and_abi import _per_shot_case

    from dqc_simulator_b200.engine import Engine

    rho, end, singles, outcomes, ex = _per_shot_case(circuit, batch, seed=17,
                                                     make=lambda f, n, b, seed: Engine(f, n, b, seed=seed),
                                                     check=8 if circuit == "grover" else 24)
    assert len({tuple(outcomes[:, s]) for s in range(batch)}) > min(batch // 4, 2 ** outcomes.shape[0] // 2)
    for s, (r1, e1) in singles.items():
        assert np.abs(rho[s] - r1).max() < 1e-10
        assert abs(end[s] - e1) < 1e-5
    assert len(set(end.tolist())) > 1


def test_heralded_link_engine_equals_oracle_shot_by_shot():
    """The probabilistic physical layer on the CUDA engine: per-shot attempt counts (CL_GEOM) and heralded Bell
    states (CL_RANDBIT) come from the shot's Philox stream inside the tile passes -- herald times, Bell bits and the
    two-qubit states equal the oracle's in every shot; with gate + memory noise the Z fix-up (135 us, its own noise)
    exists in exactly the Psi- shots."""
    from dqc_simulator_b200 import hardware

    B00 = np.array([1, 0, 0, 1], dtype=complex) / np.sqrt(2)
    for kw in (dict(), dict(p_loss_init=0.3, visibility=0.8)):
        res = []
        for mk in (gpu, cpu):
            dqc = hardware.DQCSpec.uniform(2, num_positions=5, heralded=hardware.HeraldedLinkSpec(**kw),
                                           single_qubit_gate_error_prob=1e-3, comm_qubit_depolar_rate=50.0)
            ex, be, _ = rc.run_golden(mk, "ref_distribute_ebit.json", "dm", max_qubits=4, batch=96, dqc=dqc)
            axes = ex.axes_of([(0, "node_0"), (0, "node_1")])
            link = ex.heralded_links[0]
            res.append((ex.read_time(link["herald_time"]), np.asarray(be.read_creg(link["bell_bit"])),
                        np.array([be.reduced_dm(axes, s) for s in range(96)])))
        assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
        assert np.abs(res[0][2] - res[1][2]).max() < TOL
        assert len(set(res[0][0].tolist())) > 2 and 0 < res[0][1].sum() < 96


def test_ancilla_free_cat_lowering_on_the_engine():
    """C3 shape (QFT-12 over 4 QPUs, cat-comm, ideal hardware): the ancilla-free lowering on the CUDA engine (12-qubit
    register, no measurement, a handful of passes) against the explicit protocol on the oracle (14 qubits, 252
    sampled measurements)."""
    import os

    from dqc_simulator_b200 import executor, opstream

    ops, meta = opstream.load(os.path.join(rc.GOLDEN, "c3_qft12_4qpu_cat.json"))
    dqc = hardware.DQCSpec.uniform(4, num_positions=meta["num_positions"], num_comm_qubits=2)
    qubits = [tuple(meta["qubit_map"][str(i)]) for i in range(12)]
    o = cpu("ket", 14, 1)
    exo = executor.DQCExecutor(ops, dqc, o).run()
    g = gpu("ket", 12, 1)
    exg = executor.DQCExecutor(ops, dqc, g, ancilla_free_cat=True).run()
    assert exg.measurements == 0 and g.stats()["passes"] < 12
    for lo in (0, 4, 7):
        sub = qubits[lo:lo + 5]
        _cmp(g.reduced_dm(exg.axes_of(sub), 0), o.reduced_dm(exo.axes_of(sub), 0))
