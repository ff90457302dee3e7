"""The reference's own printed 16-digit end-to-end fidelities (real NetSquid density-matrix runs with
depolarising gate / memory noise, measurement flips and Werner ebits) reproduced digit for digit.

A printed value belongs to one trajectory of measurement outcomes; tests/ref_cases.py enumerates the forced
outcomes of the example's circuit and pins (a) that exactly one trajectory hits the printed number to a few
ulp and every other one misses it by > 1e-8, on the NumPy oracle (CPU) and (b) that the CUDA engine,
driven through the C-ABI with the same forced outcomes, returns the printed number too (GPU)."""
import itertools

import pytest

import ref_cases as rc
from oracle.numpy_engine import OracleEngine

CASES = sorted(rc.PRINTED_FIDELITIES)
ULP_TOL = 2e-15   # |ours - printed|: the printed values carry 16 significant digits


def cpu(formalism, max_qubits, batch):
    return OracleEngine(formalism, max_qubits, batch, seed=1)


def gpu(formalism, max_qubits, batch):
    from dqc_simulator_b200.engine import Engine

    return Engine(formalism, max_qubits, batch, seed=1)


@pytest.mark.parametrize("case", CASES)
def test_oracle_reproduces_printed_fidelity_on_exactly_one_trajectory(case):
    _, _, _, _, _, nmeas, printed, winner, _ = rc.PRINTED_FIDELITIES[case]
    hits = []
    for outs in itertools.product((0, 1), repeat=nmeas):
        f = rc.printed_fidelity(cpu, case, outs)
        if abs(f - printed) < ULP_TOL:
            hits.append(outs)
        else:
            assert abs(f - printed) > 1e-8, (case, outs, f)   # no near misses: the match is not an accident
    assert hits == [winner], (case, hits)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_cuda_engine_reproduces_printed_fidelity(case):
    _, _, _, _, _, _, printed, winner, _ = rc.PRINTED_FIDELITIES[case]
    f = rc.printed_fidelity(gpu, case, winner)
    assert abs(f - printed) < 1e-13, (case, f, printed)
    # and a neighbouring trajectory gives the oracle's value for it, not the printed one
    other = tuple(1 - b for b in winner)
    fo, fc = rc.printed_fidelity(gpu, case, other), rc.printed_fidelity(cpu, case, other)
    assert abs(fo - fc) < 1e-12 and abs(fo - printed) > 1e-7
