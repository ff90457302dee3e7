"""State-vector sharded over GPUs by its high-order qubits (BASELINE config 5).

One process per GPU.  A register of ``n`` logical axes is split over ``R = 2^g`` ranks: every rank
holds ``2^(n-g)`` amplitudes in its own engine (all ``n-g`` local axes live), the ``g`` *global*
axes are the bits of the rank id.  What needs communication and what does not:

* gate on local axes only ............................ local pass, no communication
* diagonal gate / phase touching a global axis ....... the rank's bit selects a sub-table: local, no comm
* controlled gate whose CONTROL is global ............ ranks with bit 0 skip, ranks with bit 1 apply: no comm
* gate that MIXES (targets) a global axis ............ swap that global axis with a free local axis:
  pairwise half-shard exchange between ranks r and r ^ (1 << j) (NCCL send/recv over NVLink), then the
  gate is local.  The axis stays local afterwards (slot map), so a QFT localises each axis about once.
* measurement ........................................ local partial probabilities -> all_reduce(sum) ->
  every rank draws the same Philox number -> local collapse.

The local backend is any object with ``apply_1q / apply_2q / apply_ctrl_1q / apply_diag / sync`` and a
``buffer()`` returning the flat complex128 torch tensor of the shard (``GpuShard`` below wraps
:class:`dqc_simulator_b200.engine.Engine`; the CPU gloo tests use a NumPy stand-in).
"""
import numpy as np

_SWAP = np.eye(4, dtype=complex)[[0, 2, 1, 3]]
_CNOT = np.eye(4, dtype=complex)[[0, 1, 3, 2]]
_X = np.array([[0, 1], [1, 0]], dtype=complex)
_PAULIS = [np.eye(2, dtype=complex), _X, np.array([[0, -1j], [1j, 0]], dtype=complex), np.diag([1.0 + 0j, -1.0])]
_PAT_SWAP, _PAT_CNOT, _PAT_DIAG = (_SWAP != 0).tobytes(), (_CNOT != 0).tobytes(), (np.eye(4) != 0).tobytes()


class GpuShard:
    """Local shard on one B200: an Engine with every local axis live from the start.  Everything that touches
    the amplitudes goes through the C-ABI (gate queue, ``dqe_probs``, ``dqe_measure``, ``dqe_peer_exchange``);
    torch only lends the process group and, for tests, a view of the buffer."""

    def __init__(self, m, device, first_rank, seed=0, lazy=True):
        from .engine import Engine

        self.m = m
        self.eng = Engine("ket", m, 1, seed=seed, device=device)
        # lazy: a local slot goes live when its qubit is first used (``set_live``), like the executor's lazy
        # INSTR_INIT on one GPU -- the passes of a circuit's first stages then run over the few live bits only
        self._first_rank, self._lazy = bool(first_rank), bool(lazy)
        self._live = set()
        self._initial_state()
        self.apply_1q = self.eng.apply_1q
        self.apply_2q = self.eng.apply_2q
        self.apply_ctrl_1q = self.eng.apply_ctrl_1q
        self.apply_diag = self.eng.apply_diag
        self.apply_kq = self.eng.apply_kq
        self._peer_ptrs = None

    def _initial_state(self):
        for a in (range(1) if self._lazy else range(self.m)):
            self.set_live(a)
        if not self._first_rank:    # |0...0> lives on rank 0 only: the other shards start empty
            self.eng.apply_diag([0], [0.0, 0.0])
        self.eng.flush()

    def reset(self):
        """Back to the state right after construction, on the same handle and peer mappings (a recorded program --
        ``Engine.record`` / ``play`` -- runs again from here)."""
        self.eng.reinit()
        self._live = set()
        self._initial_state()

    def buffer(self):
        return self.eng.state_tensor()

    def set_live(self, slot):
        if slot not in self._live:
            self._live.add(slot)
            self.eng.alloc_axis(slot)

    def relive(self, slots):
        """After an exchange: exactly `slots` (plus slot 0, the scratch axis of scalings) hold live qubits."""
        self._live = set(slots) | {0}
        self.eng.force_live(sorted(self._live))

    # ---- peer memory: partners' shards mapped through CUDA IPC (NVLink loads / stores) -----------
    def enable_peer_access(self, rank, world, group=None):
        import torch.distributed as dist

        handles = [None] * world
        dist.all_gather_object(handles, self.eng.ipc_export(), group=group)
        self._rank = rank
        self._peer_ptrs = [None if p == rank else self.eng.ipc_open(handles[p]) for p in range(world)]

    def close(self):
        """Unmap the partners' shards (``cudaIpcCloseMemHandle``) and free the engine."""
        if self._peer_ptrs:
            self.eng.sync()
            for ptr in self._peer_ptrs:
                if ptr:
                    self.eng.ipc_close(ptr)
            self._peer_ptrs = None
        self.eng.close()

    def exchange(self, rank_bits, local_axes):
        """Multi-axis global <-> local exchange with all partners in one launch, synchronised on the device
        (kernel ``k_peer_exchange``): enqueue only, no host barrier."""
        self.eng.peer_exchange(self._peer_ptrs, self._rank, rank_bits, local_axes)

    def peer_1q(self, partner, my_bit, m, ctrl_axis=-1):
        """Fused gate + transfer on a global axis: this rank updates both shards on its half of the
        local index range (kernel ``k_peer_1q``)."""
        self.eng.peer_1q(self._peer_ptrs[partner], my_bit, my_bit, m, ctrl_axis)

    def peer_swap(self, partner, my_bit, local_axis):
        """Global <-> local axis swap in place over peer memory (kernel ``k_peer_swap``; host-synchronised)."""
        self.eng.peer_swap(self._peer_ptrs[partner], my_bit, my_bit, local_axis)

    def sync(self):
        self.eng.sync()

    def scale(self, s):
        s = complex(s)
        self.eng.apply_diag([0], [s, s])        # rides in the next pass of the gate queue

    def probs(self, axis):
        """(sum |amp|^2 over axis bit = 0, over bit = 1) of the local shard (kernel ``k_probs_bit``)."""
        return self.eng.probs(axis)

    norm2_by_bit = probs

    def probs_diag(self, pairs, target):
        """Sum of the diagonal elements of a density matrix held as a ket on (rows | columns) bits, by the value of
        local bit ``target`` (kernel ``k_probs_diag``)."""
        return self.eng.probs_diag(pairs, target)

    def project(self, axis, outcome, scale, p_local):
        """Keep the ``outcome`` half of ``axis`` and rescale: the engine's forced measurement renormalises by the
        LOCAL probability, the factor that turns that into the global normalisation is returned for the
        caller's pending scalar."""
        if p_local <= 0.0:
            self.eng.apply_diag([0], [0.0, 0.0])
            return 1.0
        self.eng.measure(axis, None, int(outcome), False, 0.0)
        return float(scale) * float(np.sqrt(p_local))


class ShardedKet:
    def __init__(self, n, shard, rank, world, group=None, seed=0, peer=False, peer_gates=False):
        g = int(round(np.log2(world)))
        if 2 ** g != world:
            raise ValueError("world size must be a power of two")
        if g > n - 3:
            raise ValueError("need at least 3 local axes")
        self.n, self.g, self.m = n, g, n - g
        self.rank, self.world, self.group = rank, world, group
        self.shard = shard
        self.max_qubits = n
        self.formalism = 0
        self.batch = 1
        self.slot = list(range(n))       # logical axis -> physical slot; slots >= m are rank bits
        self.live = []
        self.exchanges = 0
        self.exchanged_bytes = 0
        self.exchange_seconds = 0.0
        self._pending = 1.0 + 0.0j      # rank-local scalar phase not yet applied (folded into the next table)
        self.peer = bool(peer)          # exchanges run as ONE kernel per rank over peer memory (k_peer_swap)
        self.peer_gates = bool(peer_gates)  # also run the first gate on a global target in place (k_peer_1q);
        #                                     measured slower than swap + local gate (twice the NVLink bytes)
        self.peer_gate_count = 0
        self.peer_seconds = 0.0
        self._peer_uses = {}            # logical axis -> mixing gates served over peer memory while global
        if self.peer:
            shard.enable_peer_access(rank, world, group)
        self._seed, self._draw = int(seed), 0
        self._mixed = []                # logical axes in the order they were last mixed (victims of an exchange)
        self.relabelled_swaps = 0
        self._cx = []                   # CNOTs held back: CX(a,b) CX(b,a) CX(a,b) is a SWAP, i.e. a relabelling

    # ------------------------------------------------------------------ slots
    def _is_global(self, a):
        return self.slot[a] >= self.m

    def _rank_bit(self, a):
        return (self.rank >> (self.slot[a] - self.m)) & 1

    def _axis_at(self, slot):
        return self.slot.index(slot)

    def _touch(self, *axes):
        for a in axes:
            if a in self._mixed:
                self._mixed.remove(a)
            self._mixed.append(a)

    def _localise_all(self, busy):
        """Peer mode: ALL global axes trade places with local slots in one device-synchronised exchange
        (``dqe_peer_exchange``): (1 - 2^-g) of the shard crosses NVLink once, where g pairwise swaps move g / 2
        of it, and nothing blocks on the host.  Victims: local axes that were already mixed (a QFT / GHZ qubit
        is mixed once and then only controls), latest first; otherwise slots from 4 upwards (16-amplitude =
        256-byte runs stay contiguous on the wire)."""
        busy_slots = {self.slot[x] for x in busy}
        cand = [self.slot[x] for x in reversed(self._mixed) if not self._is_global(x)]
        cand += [s_ for s_ in range(min(4, self.m - self.g), self.m)] + list(range(self.m))
        chosen = []
        for s_ in cand:
            if s_ not in busy_slots and s_ not in chosen:
                chosen.append(s_)
            if len(chosen) == self.g:
                break
        self._apply_pending()
        self.shard.exchange(list(range(self.g)), chosen)
        for j, L in enumerate(chosen):
            a_g, other = self._axis_at(self.m + j), self._axis_at(L)
            self.slot[other], self.slot[a_g] = self.m + j, L
        self._relive()
        self.exchanges += 1
        self.exchanged_bytes += (2 ** (self.m - self.g)) * (self.world - 1) * 16

    def _localise(self, a, busy):
        """Make logical axis ``a`` local by swapping its global slot with a free local slot."""
        if not self._is_global(a):
            return
        if self.peer and not self.peer_gates:
            return self._localise_all(busy)
        busy_slots = {self.slot[x] for x in busy}
        L = next(s for s in range(self.m - 1, -1, -1) if s not in busy_slots)
        j = self.slot[a] - self.m
        self._exchange(j, L)
        other = self._axis_at(L)
        self.slot[other], self.slot[a] = self.slot[a], L
        self._peer_uses[other] = 0
        self._relive()

    def _exchange(self, j, L):
        """new[rank bit j = x, local bit L = y] = old[rank bit j = y, local bit L = x]."""
        import torch
        import torch.distributed as dist

        import time

        self._apply_pending()       # the pending phase belongs to the data that is about to move
        self.shard.sync()
        if self.peer:               # one kernel per rank over peer memory: no staging, no NCCL
            dist.barrier(group=self.group)
            t0 = time.perf_counter()
            self.shard.peer_swap(self.rank ^ (1 << j), (self.rank >> j) & 1, L)
            self.shard.sync()
            dist.barrier(group=self.group)
            self.exchanges += 1
            self.exchanged_bytes += (2 ** (self.m - 1)) * 16
            self.exchange_seconds += time.perf_counter() - t0
            return
        t0 = time.perf_counter()
        buf = self.shard.buffer()
        mybit = (self.rank >> j) & 1
        partner = self.rank ^ (1 << j)
        view = buf.view(-1, 2, 2 ** L)[:, 1 - mybit, :]
        send = torch.view_as_real(view.contiguous())
        recv = torch.empty_like(send)
        ops = [dist.P2POp(dist.isend, send, partner, group=self.group),
               dist.P2POp(dist.irecv, recv, partner, group=self.group)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        view.copy_(torch.view_as_complex(recv))
        if buf.is_cuda:
            torch.cuda.synchronize()
        self.exchanges += 1
        self.exchanged_bytes += send.numel() * 8
        self.exchange_seconds += time.perf_counter() - t0

    # ------------------------------------------------------------------ register
    def alloc_axis(self, axis=-1):
        if axis < 0:
            axis = min(a for a in range(self.n) if a not in self.live)
        self.live.append(axis)
        if not self._is_global(axis) and hasattr(self.shard, "set_live"):
            self.shard.set_live(self.slot[axis])
        return axis

    def _relive(self):
        """The engine's live slots = the local slots whose logical axis is live (after axes changed slots)."""
        if hasattr(self.shard, "relive"):
            want = {self.slot[a] for a in self.live if not self._is_global(a)} | {0}
            if want != self.shard._live:
                self.shard.relive(want)

    def free_axis(self, axis):
        raise NotImplementedError("sharded registers keep every axis (no comm-qubit churn in config 5)")

    def set_predicate(self, bit):
        if bit >= 0:
            raise NotImplementedError("predicated ops need per-shot registers; sharded mode is single-shot")

    # ------------------------------------------------------------------ gates
    def _use_peer(self, axis):
        """Policy: the first gate that mixes a global axis runs over peer memory in place (cheaper than
        a swap when the axis is mixed once, e.g. a GHZ ladder); a second one means the axis is hot (a
        QFT target): swap it into a local slot and keep it there."""
        if not (self.peer and self.peer_gates and self._is_global(axis)):
            return False
        if self._peer_uses.get(axis, 0) >= 1:
            self._peer_uses[axis] = 0
            return False
        self._peer_uses[axis] = self._peer_uses.get(axis, 0) + 1
        return True

    def _peer_gate(self, axis, m, ctrl_slot=-1, active=True):
        """2x2 gate on a GLOBAL axis without moving it: both ranks of every pair run k_peer_1q on
        their half of the index range, reading / writing the partner's shard over NVLink."""
        import time

        import torch.distributed as dist

        self._apply_pending()
        self.shard.sync()
        dist.barrier(group=self.group)          # every shard quiescent before anyone writes remotely
        t0 = time.perf_counter()
        if active:
            j = self.slot[axis] - self.m
            self.shard.peer_1q(self.rank ^ (1 << j), (self.rank >> j) & 1, m, ctrl_slot)
        self.shard.sync()
        dist.barrier(group=self.group)
        self.peer_gate_count += 1
        self.peer_seconds += time.perf_counter() - t0

    def _flush_cx(self):
        if self._cx:
            held, self._cx = self._cx, []
            for c, t in held:
                self.apply_ctrl_1q(c, t, _X)

    def _cnot(self, c, t):
        """The front end lowers SWAP to three CNOTs: recognised here, the triple costs nothing (slot relabelling)."""
        k = len(self._cx)
        if k == 1 and self._cx[0] == (t, c):
            self._cx.append((c, t))
        elif k == 2 and self._cx[0] == (c, t):
            self._cx = []
            self.slot[c], self.slot[t] = self.slot[t], self.slot[c]
            self.relabelled_swaps += 1
            self._relive()
        else:
            self._flush_cx()
            self._cx = [(c, t)]

    def apply_1q(self, axis, m):
        self._flush_cx()
        m = np.asarray(m, dtype=complex).reshape(2, 2)
        if m[0, 1] == 0 and m[1, 0] == 0:
            return self.apply_diag([axis], [m[0, 0], m[1, 1]])
        if self._use_peer(axis):
            return self._peer_gate(axis, m)
        self._localise(axis, [axis])
        self._touch(axis)
        self.shard.apply_1q(self.slot[axis], m)

    def apply_ctrl_1q(self, c, t, m):
        self._flush_cx()
        m = np.asarray(m, dtype=complex).reshape(2, 2)
        if m[0, 1] == 0 and m[1, 0] == 0:
            return self.apply_diag([c, t], [1, 1, m[0, 0], m[1, 1]])
        if self._use_peer(t):
            if self._is_global(c):
                return self._peer_gate(t, m, active=bool(self._rank_bit(c)))
            return self._peer_gate(t, m, ctrl_slot=self.slot[c])
        self._localise(t, [c, t])
        self._touch(t)
        if self._is_global(c):
            if self._rank_bit(c):
                self.shard.apply_1q(self.slot[t], m)
        else:
            self.shard.apply_ctrl_1q(self.slot[c], self.slot[t], m)

    def apply_2q(self, a, b, m):
        m = np.asarray(m, dtype=complex).reshape(4, 4)
        pat = (m != 0).tobytes()
        if pat == _PAT_CNOT and m[0, 0] == 1 and m[1, 1] == 1 and m[2, 3] == 1 and m[3, 2] == 1:
            return self._cnot(a, b)
        self._flush_cx()
        if pat == _PAT_DIAG:
            return self.apply_diag([a, b], np.diag(m))
        if pat == _PAT_SWAP and m[0, 0] == 1 and m[1, 2] == 1 and m[2, 1] == 1 and m[3, 3] == 1:
            # a SWAP is a relabelling of the slot map: no pass, no exchange, whatever slots the two axes sit in
            self.slot[a], self.slot[b] = self.slot[b], self.slot[a]
            self.relabelled_swaps += 1
            self._relive()
            return
        if not m[:2, 2:].any() and not m[2:, :2].any() and m[0, 0] == 1 and m[1, 1] == 1 and m[0, 1] == 0 and m[1, 0] == 0:
            return self.apply_ctrl_1q(a, b, m[2:, 2:])
        self._localise(a, [a, b])
        self._localise(b, [a, b])
        self._touch(a, b)
        self.shard.apply_2q(self.slot[a], self.slot[b], m)

    def apply_kq(self, axes, m):
        """Dense k-qubit matrix (k <= 4, first axis = most significant index): every axis it touches is made local."""
        self._flush_cx()
        axes = list(axes)
        for a in axes:
            self._localise(a, axes)
        self._touch(*axes)
        self.shard.apply_kq([self.slot[a] for a in axes], m)

    def apply_diag(self, axes, d):
        self._flush_cx()
        axes = list(axes)
        k = len(axes)
        d = np.asarray(d, dtype=complex).reshape((2,) * k)
        local_axes = []
        index = []
        for a in axes:
            if self._is_global(a):
                index.append(self._rank_bit(a))
            else:
                index.append(slice(None))
                local_axes.append(self.slot[a])
        sub = d[tuple(index)]
        if not local_axes:
            self._pending *= complex(sub)     # pure rank-dependent phase: no pass over the shard
            return
        table = np.asarray(sub).reshape(-1)
        if self._pending != 1:
            table = table * self._pending
            self._pending = 1.0 + 0.0j
        self.shard.apply_diag(local_axes, table)

    def depolarize(self, axes, p):
        if p:
            raise NotImplementedError("sharded mode is the ideal state-vector configuration")

    # ------------------------------------------------------------------ measurement
    def measure(self, axis, creg_bit=None, forced=-1, discard=False, flip_prob=0.0):
        """Z-basis measurement across shards; returns the outcome (identical on every rank): local partial
        probabilities (engine kernel) -> one all-reduce of two doubles -> every rank draws the same Philox
        number -> local projection (engine), the global renormalisation rides in the pending scalar."""
        import torch
        import torch.distributed as dist

        from .philox_host import shot_uniforms

        if discard or flip_prob:
            raise NotImplementedError
        self._flush_cx()
        glob = self._is_global(axis)
        q = list(self.shard.probs(0 if glob else self.slot[axis]))
        w = abs(self._pending) ** 2            # the pending scalar has not reached the amplitudes yet
        if glob:
            tot = (q[0] + q[1]) * w
            p = [tot, 0.0] if self._rank_bit(axis) == 0 else [0.0, tot]
        else:
            p = [q[0] * w, q[1] * w]
        buf = self.shard.buffer()
        t = torch.tensor(p, dtype=torch.float64, device=buf.device if buf.is_cuda else "cpu")
        dist.all_reduce(t, group=self.group)
        p0, p1 = (float(x) for x in t.cpu())
        u, _ = shot_uniforms(self._seed, 0, self._draw)
        self._draw += 1
        m = int(forced) if forced is not None and forced >= 0 else int(u * (p0 + p1) < p1)
        pm = p1 if m else p0
        scale = 1.0 / np.sqrt(pm) if pm > 0 else 0.0
        if glob:
            self._pending *= scale if self._rank_bit(axis) == m else 0.0
        else:
            self._pending *= self.shard.project(self.slot[axis], m, scale, q[m])
        return m

    # ------------------------------------------------------------------ read-out
    def _apply_pending(self):
        if self._pending != 1:
            self.shard.scale(self._pending)
            self._pending = 1.0 + 0.0j

    def flush(self):
        self._flush_cx()
        self._apply_pending()
        self.shard.sync()

    def sync(self):
        self.flush()

    def norm2(self):
        import torch
        import torch.distributed as dist

        self._flush_cx()
        q = self.shard.probs(0)
        buf = self.shard.buffer()
        t = torch.tensor([(q[0] + q[1]) * abs(self._pending) ** 2], dtype=torch.float64,
                         device=buf.device if buf.is_cuda else "cpu")
        dist.all_reduce(t, group=self.group)
        return float(t[0])

    def exchange_stats(self):
        """Seconds spent in exchanges (device time of ``k_peer_exchange`` in peer mode, host clock otherwise) and
        the error word of the device-side handshake (1 = a partner never arrived)."""
        if self.peer and not self.peer_gates and hasattr(self.shard, "eng"):
            st = self.shard.eng.peer_stats()
            return dict(seconds=st["ms"] * 1e-3 + self.exchange_seconds, error=st["error"])
        return dict(seconds=self.exchange_seconds, error=0)

    def logical_index_of(self, local_index):
        """Logical basis-state index (axis a = bit a) of local amplitude ``local_index`` on this rank."""
        phys = local_index | (self.rank << self.m)
        out = 0
        for a in range(self.n):
            out |= ((phys >> self.slot[a]) & 1) << a
        return out

    def gather_logical(self):
        """Full state in logical order on every rank (tests only: 2^n amplitudes)."""
        import torch
        import torch.distributed as dist

        self.flush()
        loc = torch.view_as_real(self.shard.buffer().clone()).contiguous()
        parts = [torch.empty_like(loc) for _ in range(self.world)]
        dist.all_gather(parts, loc, group=self.group)
        phys = torch.view_as_complex(torch.cat(parts)).cpu().numpy()     # index = rank << m | local
        perm = [self.n - 1 - self.slot[a] for a in range(self.n - 1, -1, -1)]
        return phys.reshape((2,) * self.n).transpose(perm).reshape(-1)


class ShardedDM:
    """ONE density matrix sharded over GPUs by its high-order ROW bits (third row: the C4 variant
    whose rho exceeds one GPU).  rho on n qubits is held as a ket on 2n bits -- logical axis a = column bit of
    qubit a, n + a = its row bit, exactly the engine's own dm layout -- and that ket is a :class:`ShardedKet`: the
    top log2(world) row bits are the rank id.  Everything a density-matrix backend needs is a ket operation on it:

    * U rho U^dagger ............ U on the row bit, conj(U) on the column bit
    * diagonal gates ............ table on the row bits, conjugate table on the column bits (rank-dependent
                                  sub-tables where a row bit is global: no communication)
    * one-qubit channels ........ their 4x4 superoperator as a two-"qubit" gate on (row bit, column bit)
    * two-qubit depolarising .... its 16x16 superoperator as a dense 4-axis block (OP_KBLOCK)
    * measurement ............... local sum of diagonal elements (``dqe_probs_diag``) -> all-reduce -> one Philox
                                  draw -> projection of the row AND the column bit, renormalised by 1 / p
    * predicated gates .......... the outcome is one bit for the whole matrix (one trajectory per sharded rho):
                                  ``set_predicate`` turns the following ops on or off on the host

    A gate that mixes a global row bit first exchanges the global axes with local ones (``ShardedKet._localise``:
    one device-synchronised peer kernel); column bits start local and stay local unless picked as exchange victims.
    A 17-qubit noisy rho (256 GiB) then runs on 8 x 32 GiB shards."""

    formalism = 1

    def __init__(self, n, shard, rank, world, group=None, seed=0, peer=False):
        self.n = n
        self.max_qubits = n
        self.batch = 1
        self.ket = ShardedKet(2 * n, shard, rank, world, group=group, seed=seed, peer=peer)
        self.live = []
        self._creg = 0
        self._pred = -1
        self._seed, self._draw = int(seed), 0

    # ---- register ---------------------------------------------------------------------------------------------
    def alloc_axis(self, axis=-1):
        if axis < 0:
            axis = min(a for a in range(self.n) if a not in self.live)
        self.live.append(axis)
        self.ket.alloc_axis(axis)
        self.ket.alloc_axis(self.n + axis)
        return axis

    def set_predicate(self, bit):
        self._pred = int(bit)

    def _on(self):
        return self._pred < 0 or (self._creg >> self._pred) & 1

    # ---- gates ------------------------------------------------------------------------------------------------
    def apply_1q(self, axis, m):
        if not self._on():
            return
        m = np.asarray(m, dtype=complex).reshape(2, 2)
        self.ket.apply_1q(self.n + axis, m)
        self.ket.apply_1q(axis, m.conj())

    def apply_2q(self, a, b, m):
        if not self._on():
            return
        m = np.asarray(m, dtype=complex).reshape(4, 4)
        self.ket.apply_2q(self.n + a, self.n + b, m)
        self.ket.apply_2q(a, b, m.conj())

    def apply_ctrl_1q(self, c, t, m):
        if not self._on():
            return
        m = np.asarray(m, dtype=complex).reshape(2, 2)
        self.ket.apply_ctrl_1q(self.n + c, self.n + t, m)
        self.ket.apply_ctrl_1q(c, t, m.conj())

    def apply_diag(self, axes, d):
        if not self._on():
            return
        d = np.asarray(d, dtype=complex).reshape(-1)
        self.ket.apply_diag([self.n + a for a in axes], d)
        self.ket.apply_diag(list(axes), d.conj())

    # ---- channels ---------------------------------------------------------------------------------------------
    def _superop_1q(self, axis, kraus):
        """sum_i K_i rho K_i^dagger on one qubit: S[(r, c), (r', c')] = sum_i K_i[r, r'] conj(K_i[c, c'])."""
        S = sum(np.kron(K, K.conj()) for K in (np.asarray(k, dtype=complex).reshape(2, 2) for k in kraus))
        self.ket.apply_2q(self.n + axis, axis, S)

    def pauli_channel(self, axis, p):
        if not self._on():
            return
        self._superop_1q(axis, [np.sqrt(w) * P for w, P in zip(p, _PAULIS) if w > 0])

    def kraus_1q(self, axis, ks):
        if self._on():
            self._superop_1q(axis, ks)

    def depolarize(self, axes, p):
        """(1 - p) rho + p (I / 2^k (x) Tr_axes rho), k = 1, 2 (noise_models.py:343-366)."""
        if not self._on() or not p:
            return
        axes = list(axes)
        if len(axes) == 1:
            return self.pauli_channel(axes[0], [1 - 0.75 * p, 0.25 * p, 0.25 * p, 0.25 * p])
        a, b = axes
        S = np.zeros((16, 16), dtype=complex)
        for i, Pa in enumerate(_PAULIS):
            for j, Pb in enumerate(_PAULIS):
                w = (1 - 15 * p / 16) if i == 0 and j == 0 else p / 16
                P = np.kron(Pa, Pb)
                S += w * np.kron(P, P.conj())          # index order (ra, rb, ca, cb)
        self.ket.apply_kq([self.n + a, self.n + b, a, b], S)

    # ---- measurement ------------------------------------------------------------------------------------------
    def _diag_pairs(self, skip=None):
        """(local bit, partner local bit, fixed) constraints that make a shard element a diagonal element of rho."""
        k, pairs = self.ket, []
        for q in self.live:
            r, c = k.slot[self.n + q], k.slot[q]
            rg, cg = r >= k.m, c >= k.m
            if rg and cg:
                if ((k.rank >> (r - k.m)) ^ (k.rank >> (c - k.m))) & 1:
                    return None                        # this shard holds no diagonal element at all
            elif rg:
                pairs.append((c, 0, (k.rank >> (r - k.m)) & 1))
            elif cg:
                pairs.append((r, 0, (k.rank >> (c - k.m)) & 1))
            else:
                pairs.append((r, c, -1))
        return pairs

    def measure(self, axis, creg_bit=None, forced=-1, discard=False, flip_prob=0.0):
        import torch
        import torch.distributed as dist

        from .philox_host import shot_uniforms

        if discard:
            raise NotImplementedError("sharded density matrices keep every axis")
        k = self.ket
        k._flush_cx()
        pairs = self._diag_pairs()
        r, c = k.slot[self.n + axis], k.slot[axis]
        if pairs is None:
            p = [0.0, 0.0]
        else:
            target = c if c < k.m else (r if r < k.m else -1)
            q0, q1 = k.shard.probs_diag(pairs, target)
            s = complex(k._pending)                   # rho carries the pending scalar linearly
            q0, q1 = (q0 * s).real, (q1 * s).real
            if target < 0:                             # the axis' bits are both rank bits: outcome fixed by the rank
                v = (k.rank >> (c - k.m)) & 1
                p = [q0 + q1, 0.0] if v == 0 else [0.0, q0 + q1]
            else:
                p = [q0, q1]
        buf = k.shard.buffer()
        t = torch.tensor(p, dtype=torch.float64, device=buf.device if buf.is_cuda else "cpu")
        dist.all_reduce(t, group=k.group)
        p0, p1 = (float(x) for x in t.cpu())
        e = float(flip_prob)
        q0, q1 = (1 - e) * p0 + e * p1, (1 - e) * p1 + e * p0       # MeasurementNoiseModel folded in
        u, _ = shot_uniforms(self._seed, 0, self._draw)
        self._draw += 1
        m = int(forced) if forced is not None and forced >= 0 else int(u * (q0 + q1) < q1)
        if e:
            # rho' = [(1 - e) P_m rho P_m + e P_m X rho X P_m] / q_m, kept in the |m><m| block
            raise NotImplementedError("sharded density matrices: measurement flips are applied by the caller as a Pauli channel")
        pm = p1 if m else p0
        proj = np.array([1.0, 0.0]) if m == 0 else np.array([0.0, 1.0])
        k.apply_diag([self.n + axis], proj)
        k.apply_diag([axis], proj * (1.0 / pm if pm > 0 else 0.0))
        if creg_bit is not None and creg_bit >= 0:
            self._creg = (self._creg & ~(1 << creg_bit)) | (m << creg_bit)
        return m

    def trace(self):
        import torch
        import torch.distributed as dist

        k = self.ket
        k._flush_cx()
        pairs = self._diag_pairs()
        q = (0.0, 0.0) if pairs is None else k.shard.probs_diag(pairs, -1)
        buf = k.shard.buffer()
        t = torch.tensor([((q[0] + q[1]) * complex(k._pending)).real], dtype=torch.float64,
                         device=buf.device if buf.is_cuda else "cpu")
        dist.all_reduce(t, group=k.group)
        return float(t[0])

    # ---- read-out ---------------------------------------------------------------------------------------------
    def flush(self):
        self.ket.flush()

    sync = flush

    @property
    def exchanges(self):
        return self.ket.exchanges

    def gather_rho(self):
        """Full 2^n x 2^n matrix on every rank (tests only)."""
        v = self.ket.gather_logical()
        return v.reshape(2 ** self.n, 2 ** self.n)
