"""State-vector sharding on REAL devices: 2 and 4 GPUs of one box, one process per GPU
(torchrun), exchanges over CUDA-IPC peer memory synchronised on the device (``k_peer_exchange``), measurement
through engine kernels.  ``gather_logical()`` against the NumPy oracle, amplitude by amplitude, at n = 14;
skipped where fewer devices are visible (the driver's single-GPU tier)."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world,peer", [(2, 1), (4, 1), (2, 0)])
def test_sharded_ket_matches_oracle_on_real_devices(world, peer, tmp_path):
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = tmp_path / "res.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "run_sharded.py"),
           "--qubits", "14", "--circuit", "random", "--peer", str(peer), "--out", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    res = json.loads(out.read_text())
    assert res["max_abs_err"] < 1e-10 and res["exchanges"] > 0 and res["relabelled_swaps"] > 0 and not res["peer_error"]


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_density_matrix_matches_oracle_on_real_devices(world, tmp_path):
    """ONE rho (n = 7: 16384 elements) sharded by its high-order row bits over real devices (ShardedDM): gates, 1- and
    2-qubit depolarising (OP_KBLOCK superoperator), Pauli / Kraus channels, forced measurements with a predicated
    correction -- element by element against the oracle's density-matrix run."""
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = tmp_path / "res.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "run_sharded.py"),
           "--qubits", "7", "--circuit", "random_dm", "--out", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    res = json.loads(out.read_text())
    assert res["max_abs_err"] < 1e-10 and abs(res["trace"] - 1) < 1e-10 and res["exchanges"] > 0


@pytest.mark.parametrize("circuit", ["qft", "ghz"])
def test_replayed_sharded_program_on_real_devices(circuit, tmp_path):
    """The second repetition of a sharded program replays the C-ABI call stream recorded during the first
    (``Engine.record`` / ``play``; passes, the device-synchronised exchange, passes) on the reset shards: same
    closed-form output (checked inside ``run_fixture``), no executor walk."""
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = tmp_path / "res.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "run_sharded.py"),
           "--qubits", "12", "--circuit", circuit, "--reps", "3", "--out", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    res = json.loads(out.read_text())
    assert res["replayed"] and res["recorded_calls"] > 10 and res["max_abs_err"] < 1e-10 and not res["peer_error"]
    assert res["host_enqueue_seconds"] < res["first_run_host_enqueue_seconds"]
