"""BASELINE.json configs at their full single-GPU sizes, checked through size-independent properties
(the oracle cannot hold these states): norm conservation, the closed-form output of QFT|0..0> (uniform
magnitude 2^{-n/2} on the data qubits, everything else exactly zero), GHZ amplitudes, and the invariance
of the result under the planner's fusion switches."""
import os

import numpy as np
import pytest

from dqc_simulator_b200 import executor, hardware, opstream

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _run(name, nq, nqpu, npos, **opts):
    import torch

    from dqc_simulator_b200.engine import Engine

    ops, meta = opstream.load(os.path.join(GOLDEN, name))
    e = Engine("ket", nq, 1, seed=11)
    for k, v in opts.items():
        e.set_option(k, v)
    dqc = hardware.DQCSpec.uniform(nqpu, num_positions=npos, num_comm_qubits=2 if nqpu > 1 else 0)
    ex = executor.DQCExecutor(ops, dqc, e).run()
    e.sync()
    return e, ex, meta, torch


def test_config3_qft26_over_4_qpus_full_size():
    """C3: MQT-bench-shaped QFT-26 over 4 QPUs (542 cat-comm remote gates, 1084 sampled measurements),
    28-axis register = 4 GiB.  QFT|0> is the uniform superposition whatever the sampled outcomes."""
    e, ex, meta, torch = _run("c3_qft26_4qpu_cat.json.gz", 28, 4, 9)
    assert ex.finished and ex.measurements == 1084 and ex.ebits == 542
    st = e.state_tensor()
    mag = st.abs()
    assert abs(float((mag * mag).sum()) - 1.0) < 1e-9
    nz = mag > 0
    assert int(nz.sum()) == 2 ** 26                       # support = the 26 data axes, comm axes back in |0>
    assert float((mag[nz] - 2.0 ** -13).abs().max()) < 1e-10
    assert len(e.live) == 26
    stats = e.stats()
    assert stats["passes"] < stats["api_ops"] / 4          # the flusher fused the stream


def test_config5_shape_ghz30_and_qft30_single_gpu():
    e, ex, meta, torch = _run("c5_ghz30_mono.json.gz", 30, 1, 30)
    st = e.state_tensor()
    assert abs(abs(complex(st[0])) - 2 ** -0.5) < 1e-12 and abs(abs(complex(st[2 ** 30 - 1])) - 2 ** -0.5) < 1e-12
    assert int((st.abs() > 0).sum()) == 2
    e.close()
    del st, e
    torch.cuda.empty_cache()
    e, ex, meta, torch = _run("c5_qft30_mono.json.gz", 30, 1, 30)
    mag = e.state_tensor().abs()
    assert float((mag - 2.0 ** -15).abs().max()) < 1e-10
    assert e.stats()["passes"] < 120


def test_fusion_switches_do_not_change_the_state():
    """Same circuit, same sampled outcomes, planner with and without peephole / pass fusion: the states
    agree to 1e-12 (16M amplitudes)."""
    ref = None
    for opts in ({}, {"merge": 0}, {"fuse": 0}, {"tile_bits": 9}, {"tma": 0}, {"persistent": 1},
                 {"threads": 256, "tile_bits": 12}, {"cluster": 0}, {"tile_bits": 12, "cluster": 1}):
        e, ex, meta, torch = _run("c3_qft12_4qpu_cat.json", 14, 4, 5, **opts)
        st = e.state_tensor().clone()
        creg = e.read_creg_words()
        if ref is None:
            ref = (st, creg)
        else:
            assert np.array_equal(creg, ref[1]), opts
            assert float((st - ref[0]).abs().max()) < 1e-12, opts
        e.close()


def test_config2_grover5_full_batch_properties():
    """C2 at BASELINE's batch (1e5 noisy density-matrix shots of Grover-5 over 3 QPUs, 26 GB): every fidelity is a
    probability, the batch mean sits at the reference's printed trajectory value (examples/run_experiment.py:91) within
    the trajectory spread, the comm axes are back in |0> (free), and the first shots equal -- classical registers bit
    for bit, fidelities to 1e-12 -- a 64-shot engine seeded alike (the Philox stream is keyed by the global shot id,
    so neither the batch size nor a rank's share of it changes a shot)."""
    from dqc_simulator_b200 import experiment
    from dqc_simulator_b200.engine import Engine

    ops, _ = experiment.load_c2("grover")
    dqc = experiment.c2_hardware(noisy=True)
    data, ideal = experiment.ideal_output_ket(ops, lambda f, n, b: Engine(f, n, b))
    big = Engine("dm", 7, 100000, seed=2024, shot_offset=0)
    rb = experiment.run_shot_batch(ops, dqc, big, data, ideal)
    assert rb.fidelity.shape == (100000,)
    assert float(rb.fidelity.min()) > -1e-12 and float(rb.fidelity.max()) < 1 + 1e-12
    assert abs(float(rb.fidelity.mean()) - 0.1071574284590638) < 2e-5
    assert len(big.live) == 5
    # (run_shot_batch starts with reinit, which moves a handle on by one batch of shot ids: both then start at 100000)
    small = Engine("dm", 7, 64, seed=2024, shot_offset=100000 - 64)
    rs = experiment.run_shot_batch(ops, dqc, small, data, ideal)
    assert np.array_equal(rb.creg[:64], rs.creg)
    assert float(np.abs(rb.fidelity[:64] - rs.fidelity).max()) < 1e-12
    for shot in (0, 63, 99999):   # Tr rho = 1, rho Hermitian with purity <= 1, shot by shot
        r = big.reduced_dm(rb.executor.axes_of(data), shot)
        assert abs(np.trace(r) - 1) < 1e-12 and np.abs(r - r.conj().T).max() < 1e-12
        assert 0 < np.real(np.trace(r @ r)) <= 1 + 1e-12
    big.close(); small.close()


@pytest.mark.parametrize("n,shots", [(6, 8192), (16, 6144)])
def test_config4_noisy_qft_dm_equals_the_ket_trajectory_average(n, shots):
    """C4: the noisy QFT density matrix (n = 16: 4^16 complex128 = 64 GiB, BASELINE's size) against an INDEPENDENT code
    path of the engine -- the same circuit and hardware on state VECTORS, where every channel is a sampled Pauli
    (one noise trajectory per shot): <phi| rho_S |phi> of the exact density matrix must equal the mean of
    |<phi|psi_s>|^2-type overlaps over the trajectories within 5 standard errors, for product and entangled probe
    states phi on three of the qubits; and the noise must be visible at that resolution (the noiseless circuit's values
    would fail: the test has power).
    Also Tr rho = 1 and purity < 1."""
    from dqc_simulator_b200 import experiment
    from dqc_simulator_b200.engine import Engine

    name = f"c4_qft{n}_mono.json" + (".gz" if n > 12 else "")
    ops, _ = opstream.load(os.path.join(GOLDEN, name))
    dqc = hardware.DQCSpec.uniform(1, num_positions=n, num_comm_qubits=0, **experiment.C2_NOISE)
    rng = np.random.default_rng(5)
    plus = np.full(8, 8 ** -0.5, dtype=complex)
    ghz = np.zeros(8, dtype=complex); ghz[0] = ghz[7] = 2 ** -0.5
    rnd = rng.normal(size=8) + 1j * rng.normal(size=8); rnd /= np.linalg.norm(rnd)
    zero = np.zeros(8, dtype=complex); zero[0] = 1
    probes = [plus, ghz, rnd, zero]

    e = Engine("dm", n, 1, seed=3)
    ex = executor.DQCExecutor(ops, dqc, e).run()
    axes = ex.axes_of(ex.processing_qubits())
    pick = [axes[0], axes[n // 2], axes[-1]]
    exact = [e.fidelity_ket(pick, p) for p in probes]
    r3 = e.reduced_dm(pick, 0)
    assert abs(np.trace(r3) - 1) < 1e-10
    assert 0.125 - 1e-9 <= np.real(np.trace(r3 @ r3)) < 1 - 1e-6          # mixed: the channels did something
    for p, f in zip(probes, exact):
        assert abs(f - np.real(p.conj() @ r3 @ p)) < 1e-10                # the two read-out kernels agree
    e.close()
    del e

    i = Engine("ket", n, 1, seed=1)                                         # the noiseless circuit, for the power check
    exi = executor.DQCExecutor(ops, hardware.DQCSpec.uniform(1, num_positions=n, num_comm_qubits=0), i).run()
    axi = exi.axes_of(exi.processing_qubits())
    noiseless = [i.fidelity_ket([axi[0], axi[n // 2], axi[-1]], p) for p in probes]
    i.close()

    k = Engine("ket", n, shots, seed=17)
    exk = executor.DQCExecutor(ops, dqc, k).run()
    axk = exk.axes_of(exk.processing_qubits())
    pk = [axk[0], axk[n // 2], axk[-1]]
    ideal_gap = 0.0
    for p, f, f0 in zip(probes, exact, noiseless):
        per_shot = k.fidelity_ket_all(pk, p)
        se = max(float(per_shot.std(ddof=1)) / np.sqrt(shots), 1e-6)
        assert abs(float(per_shot.mean()) - f) < 5 * se, (f, float(per_shot.mean()), se)
        ideal_gap = max(ideal_gap, abs(f - f0) / se)
    assert ideal_gap > 8      # the noiseless answer would fail the same comparison
    k.close()
