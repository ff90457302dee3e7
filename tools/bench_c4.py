"""BASELINE config 4 on one GPU: noisy QFT-16 density matrix (4^16 complex128 = 64 GiB), one
trajectory, channels as config 2 (2q / 1q depolarising, memory depolarising).  Reports wall time,
passes, algorithmic GB/s; checks Tr(rho) = 1 and 0 < purity <= 1 through reduced density matrices.
Multi-GPU: one trajectory per rank (shot sharding, see bench.py) -- run N independent copies."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dqc_simulator_b200 import executor, experiment, hardware, opstream  # noqa: E402
from dqc_simulator_b200.engine import Engine  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    if world > 1:   # one trajectory per GPU (global shot id = rank): no data-path collective
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    name = f"c4_qft{n}_mono.json" + (".gz" if n > 12 else "")
    ops, _ = opstream.load(os.path.join(GOLDEN, name))
    dqc = hardware.DQCSpec.uniform(1, num_positions=n, num_comm_qubits=0, **experiment.C2_NOISE)
    e = Engine("dm", n, 1, seed=3, shot_offset=rank, device=local)
    out = None
    for rep in range(2):
        e.reinit(); e.reset_stats(); e.sync()
        e.time_passes(True)
        if world > 1:
            dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        ex = executor.DQCExecutor(ops, dqc, e).run()
        e.sync()
        if world > 1:
            dist.barrier()
        wall = time.perf_counter() - t0
        ms, npass, amps = e.pass_time()
        e.time_passes(False)
        st = e.stats()
        out = dict(case=name, wall_s=wall, pass_ms=ms, passes=npass, api_ops=st["api_ops"], micro_ops=st["micro_ops"],
                   GBps=32.0 * amps / (ms * 1e-3) / 1e9, state_GiB=4 ** n * 16 / 2 ** 30)
    axes = ex.axes_of(ex.processing_qubits())
    r1 = e.reduced_dm(axes[:1], 0)
    r3 = e.reduced_dm(axes[:3], 0)
    out["trace"] = float(np.real(np.trace(r1)))
    out["purity_3q"] = float(np.real(np.trace(r3 @ r3)))
    assert abs(out["trace"] - 1) < 1e-9 and 0 < out["purity_3q"] <= 1 + 1e-9
    out["n_gpus"] = world
    out["trajectories_per_s"] = world / out["wall_s"]
    if world > 1:
        t = torch.tensor([out["wall_s"]], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["wall_s"] = float(t[0])
        out["trajectories_per_s"] = world / out["wall_s"]
    if rank == 0:
        print(json.dumps(out), flush=True)
        os.makedirs("gpurun_out", exist_ok=True)
        with open(f"gpurun_out/bench_c4_qft{n}_x{world}.json", "w") as f:
            json.dump(out, f)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
