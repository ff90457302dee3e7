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
