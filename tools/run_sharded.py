"""BASELINE config 5: GHZ-n / QFT-n state vector sharded over N GPUs by its high-order qubits.
Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P tools/run_sharded.py --qubits 30 [--circuit qft|ghz]
Self-checking (norm, closed-form amplitudes); rank 0 prints one JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from dqc_simulator_b200 import executor, hardware, opstream  # noqa: E402
from dqc_simulator_b200.sharded import GpuShard, ShardedKet  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", dest="n", type=int, default=30)
    ap.add_argument("--circuit", default="qft")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--peer", type=int, default=1, help="1: exchanges as one kernel per rank over CUDA-IPC peer memory (k_peer_swap); 2: also in-place peer gates")
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    g = int(np.log2(world))
    suffix = ".json" if a.n <= 12 else ".json.gz"
    ops, _ = opstream.load(os.path.join(GOLDEN, f"c5_{a.circuit}{a.n}_mono{suffix}"))
    dqc = hardware.DQCSpec.uniform(1, num_positions=a.n, num_comm_qubits=0)
    best = None
    for rep in range(a.reps):
        shard = GpuShard(a.n - g, local, rank == 0)
        sk = ShardedKet(a.n, shard, rank, world, peer=bool(a.peer), peer_gates=a.peer >= 2)
        shard.eng.time_passes(True)
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        ex = executor.DQCExecutor(ops, dqc, sk).run()
        t_host = time.perf_counter() - t0
        sk.sync()
        dist.barrier(); torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        buf = shard.buffer()
        mag = buf.abs()
        norm = sk.norm2()
        if a.circuit == "qft":
            err = float((mag - 2.0 ** (-a.n / 2)).abs().max())
        else:
            err = 0.0
            want = {0: 2 ** -0.5, 2 ** a.n - 1: 2 ** -0.5}
            nz = torch.nonzero(mag > 0).flatten().tolist()
            for i in nz:
                err = max(err, abs(float(mag[i]) - want.get(sk.logical_index_of(i), 0.0)))
            cnt = torch.tensor([len(nz)], device=buf.device)
            dist.all_reduce(cnt)
            err = max(err, abs(int(cnt[0]) - 2))
        t = torch.tensor([err, dt, sk.exchange_seconds], dtype=torch.float64, device=buf.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        pass_ms, n_pass, amps = shard.eng.pass_time()
        st = shard.eng.stats()
        res = dict(circuit=a.circuit, n=a.n, n_gpus=world, seconds=float(t[1]), max_abs_err=float(t[0]),
                   norm=norm, exchanges=sk.exchanges, exchange_bytes_per_rank=sk.exchanged_bytes,
                   exchange_seconds=float(t[2]), peer=bool(a.peer), peer_gates=sk.peer_gate_count, peer_seconds=sk.peer_seconds,
                   host_enqueue_seconds=t_host, pass_ms=pass_ms, timed_passes=n_pass,
                   micro_ops=st["micro_ops"], pass_GBps=32.0 * amps / max(pass_ms, 1e-9) / 1e6,
                   nvlink_GBps_per_dir_per_gpu=(sk.exchanged_bytes / float(t[2]) / 1e9) if sk.exchanges else None,
                   passes_per_rank=st["passes"], api_ops=st["api_ops"], shard_bytes=(2 ** (a.n - g)) * 16,
                   gate_updates_per_s=len([1 for _ in range(1)]) and (sum(len(s) for sl in ops.values() for s in sl) / float(t[1])))
        assert abs(norm - 1) < 1e-9 and res["max_abs_err"] < 1e-10, res
        if best is None or res["seconds"] < best["seconds"]:
            best = res
        shard.eng.close()
        del shard, sk, buf, mag
        torch.cuda.empty_cache()
    if rank == 0:
        print(json.dumps(best), flush=True)
        os.makedirs("gpurun_out", exist_ok=True)
        with open(f"gpurun_out/sharded_{a.circuit}{a.n}_x{world}{'_peer' if a.peer else ''}.json", "w") as f:
            json.dump(best, f)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
