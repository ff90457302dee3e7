"""Executor: walks the compiled per-QPU op lists and drives a state backend.

This is the host side of the drop-in boundary.  It consumes exactly what the reference's
compilers emit -- ``qpu_op_dict = {node_name: [[op, ...] per time slice]}``
(``software/compilers.py:665-792``) -- and reproduces the control flow of
``InterpreterProtocol`` / ``DQCMasterProtocol`` (``software/dqc_control.py:899-949,
1170-1202``) without NetSquid's discrete-event kernel:

* :class:`QuantumProgram` mirrors ``QuantumProgram.apply(instruction, qubit_indices,
  operator=None, output_key="last")`` and ``program.output[key]``
  (``dqc_control.py:191-196, 307, 371, 464-465, 787``).
* :meth:`QPU.execute_program` is the level-2 boundary: it plays the
  role of ``QuantumProcessor.execute_program`` -- consume the instruction's duration, apply
  memory noise for the idle time of every touched position, apply the instruction and its
  attached noise model in the order fixed by ``hardware/quantum_processors.py:344-456`` --
  but instead of NumPy arithmetic it *enqueues* ops on the backend (the CUDA engine's gate
  queue; flushed into fused, batched launches by the C library).
* Remote-gate roles: cat ``entangle / correct / disentangle_start / disentangle_end``
  (``dqc_control.py:266-403``), tp ``bsm / correct / correct4tele_only /
  swap_then_correct / correct_manually`` (``:405-613, 686-750``), ``logged``
  (``:754-791``), ``distribute_ebit`` (``:803-832``).

Batched shots: measurement results never come back to the host.  A measurement writes a
per-shot classical bit on the device (:class:`CBit`); the classically controlled X / Z
corrections become per-shot *predicated* gates, and every shot follows ITS OWN simulated
timeline like a reference trajectory does: once a time depends on an outcome it lives in a
per-shot classical register of the backend (:mod:`.timeline`), a predicated instruction
advances the clocks of exactly the shots that run it, ``_cat_disentangle_end`` executes the
pending program in exactly the shots whose outcome was 1 (``execute_program(when=bit)``),
and memory-noise channels take their probability from a register.  With
``host_branching=True`` (batch 1 only) or forced outcomes the executor instead knows each
outcome on the host and branches exactly like the reference; all times are then floats.

Qubits are materialised lazily: ``INSTR_INIT`` only records the time at which the memory-noise
clock of a position starts (the reference's tests pin "decoherence starts after
initialisation", ``tests/unit_tests/hardware/test_noise_model.py:913-993``); an engine axis
is claimed on first use.  An untouched |0> qubit is a product factor, so this is exact and
keeps the joint register small (SURVEY.md section 7 "comm-qubit axes blow up memory").
"""
import math

import numpy as np

from . import instructions as ins
from . import timeline as tlm
from .hardware import DQCSpec


class UnfinishedQuantumCircuitError(Exception):
    """Same name and role as ``dqc_control.py:25-30``."""


class CBit:
    """Handle to a per-shot classical bit held on the device."""

    __slots__ = ("index",)

    def __init__(self, index):
        self.index = index

    def __repr__(self):
        return f"CBit({self.index})"


class _Instr:
    """One pending instruction of a :class:`QuantumProgram`.  ``done`` is the classical bit marking the
    shots in which an earlier conditional execution (``execute_program(when=bit)``) already ran it."""

    __slots__ = ("instruction", "positions", "operator", "key", "predicate", "done")

    def __init__(self, instruction, positions, operator, key, predicate):
        self.instruction, self.positions, self.operator = instruction, positions, operator
        self.key, self.predicate, self.done = key, predicate, None

    def __iter__(self):   # (instruction, positions, operator, key, predicate), the shape tests unpack
        return iter((self.instruction, self.positions, self.operator, self.key, self.predicate))


class QuantumProgram:
    """List of pending instructions; ``output[key]`` is filled by ``execute_program``."""

    def __init__(self):
        self.instructions = []
        self.output = {}

    def apply(self, instruction, qubit_indices, operator=None, output_key="last", predicate=None):
        if isinstance(qubit_indices, (int, np.integer)):
            qubit_indices = [int(qubit_indices)]
        else:
            qubit_indices = [int(q) for q in qubit_indices]
        self.instructions.append(_Instr(instruction, qubit_indices, operator, output_key, predicate))

    def __len__(self):
        return len(self.instructions)


class _CregAllocator:
    """Classical bits of the per-shot register word.  A released bit waits until the next measurement
    before it is handed out again: the backend may run classical instructions earlier than the quantum
    ops enqueued between them (see :mod:`.timeline`), so a bit must not be rewritten while an op enqueued
    earlier in the same span is still predicated on it."""

    def __init__(self, nbits):
        self.nbits = nbits
        self.free = list(range(nbits))
        self.limbo = []
        self.peak = 0

    def alloc(self):
        if not self.free:
            raise RuntimeError("out of classical bits; create the engine with more creg bits")
        b = self.free.pop(0)
        self.peak = max(self.peak, self.nbits - len(self.free))
        return b

    def release(self, b):
        if b not in self.limbo and b not in self.free:
            self.limbo.append(b)

    def fence(self):
        if self.limbo:
            self.free = sorted(self.free + self.limbo)
            self.limbo = []


class QPU:
    """Per-node memory bookkeeping + ``execute_program``."""

    def __init__(self, spec, executor):
        self.spec = spec
        self.name = spec.name
        self.ex = executor
        self.comm_qubit_positions = spec.comm_qubit_positions
        self.processing_qubit_positions = spec.processing_qubit_positions
        self.comm_qubits_free = list(self.comm_qubit_positions)  # quantum_processors.py:122
        self.ebit_ready = False                                   # never set True upstream
        self.clock = 0.0
        self.alias = set()     # comm positions that are aliases of a remote data qubit's axis (ancilla_free_cat)
        self.axis = {}         # pos -> engine axis (materialised qubits)
        self.pending = {}      # pos -> time the fresh |0> became valid (not yet materialised)
        self.last_access = {}  # pos -> time of last access (materialised qubits)

    # ---- qubit materialisation -------------------------------------------------
    def occupied(self, pos):
        return pos in self.axis or pos in self.pending

    def _materialise(self, pos):
        if pos in self.axis:
            return self.axis[pos]
        if pos not in self.pending:
            raise RuntimeError(f"{self.name}: memory position {pos} holds no qubit")
        ax = self.ex._claim_axis(pos in self.comm_qubit_positions)
        self.ex.backend.alloc_axis(ax)
        self.axis[pos] = ax
        self.last_access[pos] = self.pending.pop(pos)
        return ax

    def _drop(self, pos):
        """Throw away whatever sits at ``pos``."""
        if pos in self.alias:                 # an alias of another QPU's data qubit: nothing of its own to free
            self.alias.discard(pos)
            self.axis.pop(pos, None)
            self.last_access.pop(pos, None)
            self.pending.pop(pos, None)
            return
        if pos in self.axis:
            ax = self.axis.pop(pos)
            self.ex.backend.free_axis(ax)
            self.ex._release_axis(ax)
            self.last_access.pop(pos, None)
        self.pending.pop(pos, None)

    def _memory_noise(self, pos, now):
        """Idle-time channel on access: p = 1 - exp(-dt * 1e-9 * rate)
        (``noise_models.py:248-253``; NetSquid ``DepolarNoiseModel`` time dependent).  With per-shot
        times the exponent is evaluated per shot by the backend (``dyn_channel``)."""
        last = self.last_access[pos]
        be, ax, sp, T = self.ex.backend, self.axis[pos], self.spec, self.ex.T
        if T.same(now, last):
            return
        r_depol, r_dephase, r_amp, r_zflip = sp.mem_rates(pos)
        # (amplitude damping first, then the dephasing part: the order of NetSquid's T1T2NoiseModel [NS-doc])
        rates = ((tlm.DYN_DEPOL, r_depol), (tlm.DYN_AMPDAMP, r_amp), (tlm.DYN_DEPHASE, r_dephase),
                 (tlm.DYN_ZFLIP, r_zflip))
        if not any(r > 0 for _, r in rates):
            return
        if (T.is_reg(now) or T.is_reg(last)) and not (T.is_reg(now) and T.is_reg(last) and now.reg is last.reg):
            for kind, rate in rates:    # the idle time differs from shot to shot
                if rate > 0:
                    be.dyn_channel(ax, kind, T.qexp_between(now, last, 1e-9 * rate).index)
            return
        dt = T.sub(now, last)           # a host float: the same idle time in every shot
        if dt <= 0:
            return
        if r_depol > 0:
            be.depolarize([ax], 1.0 - math.exp(-dt * 1e-9 * r_depol))
        if r_amp > 0:
            g = 1.0 - math.exp(-dt * 1e-9 * r_amp)
            be.kraus_1q(ax, [[[1, 0], [0, math.sqrt(1 - g)]], [[0, math.sqrt(g)], [0, 0]]])
        if r_dephase > 0:
            pz = 0.5 * (1.0 - math.exp(-dt * 1e-9 * r_dephase))
            be.pauli_channel(ax, [1.0 - pz, 0.0, 0.0, pz])
        if r_zflip > 0:
            pz = 1.0 - math.exp(-dt * 1e-9 * r_zflip)
            be.pauli_channel(ax, [1.0 - pz, 0.0, 0.0, pz])

    def touch(self, positions, now):
        axes = []
        for pos in positions:
            ax = self._materialise(pos)
            self._memory_noise(pos, now)
            axes.append(ax)
        return axes

    # ---- the boundary ----------------------------------------------------------
    def execute_program(self, program, when=None):
        """Run the program's instructions on this QPU's clock, scheduled the way NetSquid's
        ``QuantumProcessor`` runs a ``QuantumProgram(parallel=True)`` (its default, which is what
        ``dqc_control.py`` builds): instructions are *issued in order*, one whose physical instruction is
        ``parallel=True`` (every gate and the measurement, ``quantum_processors.py:344-456``) starts as soon
        as the positions it acts on are free, so gates of one program on different qubits overlap in
        simulated time; ``INSTR_INIT`` / ``INSTR_DISCARD`` are ``parallel=False`` and wait for, and hold, the
        whole processor.  ``execute_program`` returns when every instruction has finished.  The start time
        of an instruction is what its memory-noise exponent is computed from.

        ``when`` (a :class:`CBit`, batched shots only): the reference reaches this call only in the
        trajectories where that bit is 1 (``_cat_disentangle_end``, ``dqc_control.py:385-403``).  The
        instructions run, and the clocks advance, in exactly those shots; the program keeps its
        instructions, marked done-in-those-shots, and the next unconditional call runs them in the
        remaining shots."""
        ex = self.ex
        be, T = ex.backend, ex.T
        overlap = ex.parallel_programs
        exact = ex.per_shot_time
        t0 = self.clock
        t_issue = t0              # in-order issue: no instruction starts before its predecessor has started
        busy = {}                 # pos -> time its running instruction finishes
        temps = []                # derived classical bits of this call

        def start_of(positions, exclusive):
            if not overlap:
                return self.clock
            if exclusive:
                return T.max(t_issue, *busy.values())
            return T.max(t_issue, *[busy[p_] for p_ in positions if p_ in busy])

        def finish(positions, start, duration, exclusive, run=None):
            nonlocal t_issue
            end = T.add(start, duration)
            if overlap:
                for p_ in positions:
                    busy[p_] = T.sel(run, end, busy.get(p_, t0))
                t_issue = T.sel(run, end if exclusive else start, t_issue)
            else:
                self.clock = T.sel(run, end, self.clock)
            return end

        for it in program.instructions:
            instruction, positions, operator, key, predicate = it
            run = predicate
            if exact and it.done is not None and it.done is predicate:
                continue          # ran wherever its predicate holds (the Z of a conditional execute_program)
            if exact:
                if it.done is not None:
                    run = ex._p_and(ex._p_not(it.done, temps), run, temps)
                run = ex._p_and(when, run, temps)
            tok, mat = ins.resolve((instruction, operator) if operator is not None else instruction)
            ph = self.spec.phys(tok)
            structural = tok.kind in ("init", "discard", "measure") or (tok is ins.INSTR_SWAP and ex.swap_as_three_cnots)
            if structural and (run is not None):
                raise NotImplementedError(
                    f"{self.name}: {tok.name} inside a conditionally executed program is not supported with "
                    "batched shots (it changes the register layout per shot); run with host_branching=True")
            if tok.kind == "init":
                # fresh |0> replaces any occupant; noise clock starts when INIT completes
                for pos in positions:
                    self._drop(pos)
                end = finish(positions, start_of(positions, True), ph.duration, True)
                for pos in positions:
                    self.pending[pos] = end
                continue
            if tok.kind == "discard":
                for pos in positions:
                    self._drop(pos)
                finish(positions, start_of(positions, True), ph.duration, True)
                continue
            if tok.kind == "measure":
                (pos,) = positions
                start = start_of(positions, False)
                (ax,) = self.touch(positions, start)
                bit = ex.creg.alloc()
                forced = ex._next_forced()
                be.measure(ax, bit, forced, ph.discard, ph.prob)
                # classical instructions enqueued from here on may be hoisted up to this measurement
                T.fence()
                ex.creg.fence()
                end = finish(positions, start, ph.duration, False)
                if ph.discard:
                    self.axis.pop(pos)
                    self.last_access.pop(pos, None)
                    ex._release_axis(ax)
                else:
                    self.last_access[pos] = end
                if forced in (0, 1):
                    # outcome fixed by the caller: known on the host for every shot of the batch, so the
                    # control flow (and with it the simulated timeline) branches exactly like the reference
                    ex.creg.release(bit)
                    program.output[key] = [forced]
                else:
                    program.output[key] = [ex._deliver(bit)]
                continue
            if tok is ins.INSTR_SWAP and ex.swap_as_three_cnots:
                # NoisyQPU composite instruction (quantum_processors.py:901-909): its own program, run after
                # everything issued so far has finished
                a, b = positions
                sub = QuantumProgram()
                sub.apply(ins.INSTR_CNOT, [a, b])
                sub.apply(ins.INSTR_CNOT, [b, a])
                sub.apply(ins.INSTR_CNOT, [a, b])
                if overlap:
                    self.clock = T.max(t_issue, *busy.values())
                self.execute_program(sub)
                t_issue = self.clock
                continue
            if self.alias and any(p_ in self.alias for p_ in positions):
                # an aliased comm qubit may only CONTROL (first operand of a controlled gate) or carry a diagonal
                m_ = np.asarray(mat)
                d_ = m_.shape[0] // 2
                ctrl_ok = (len(positions) == 2 and positions[0] in self.alias and positions[1] not in self.alias and
                           not m_[:d_, d_:].any() and not m_[d_:, :d_].any() and np.array_equal(m_[:d_, :d_], np.eye(d_)))
                if not ctrl_ok and (m_ - np.diag(np.diag(m_))).any():
                    raise NotImplementedError(f"{self.name}: ancilla_free_cat: a cat-entangled comm qubit is mixed by "
                                              f"{tok.name}; only controls and diagonal gates commute with the copy")
                ex.aliased_gates += 1
            start = start_of(positions, False)
            pred = -1 if run is None else run.index
            if pred >= 0:
                be.set_predicate(pred)
            if exact or pred < 0:
                axes = self.touch(positions, start)
            else:
                # legacy shared timeline: the idle-time channel is applied to every shot
                be.set_predicate(-1)
                axes = self.touch(positions, start)
                be.set_predicate(pred)
            if len(axes) == 1:
                be.apply_1q(axes[0], mat)
            elif len(axes) == 2:
                be.apply_2q(axes[0], axes[1], mat)
            else:
                raise NotImplementedError("only 1- and 2-qubit instructions exist upstream")
            if ph.noise in ("depol1", "depolk") and ph.prob > 0:
                be.depolarize(axes, ph.prob)
            if pred >= 0:
                be.set_predicate(-1)
            if exact:
                # a predicated gate exists, takes time and touches its qubits only in the shots that run it
                end = finish(positions, start, ph.duration, False, run)
                for pos in positions:
                    self.last_access[pos] = T.sel(run, end, self.last_access[pos])
            else:
                # legacy: one shared timeline that advances by the expected duration (see DQCExecutor)
                dur = ph.duration * (ex.predicated_time_weight if pred >= 0 else 1.0)
                end = finish(positions, start, dur, False)
                for pos in positions:
                    self.last_access[pos] = end
            program.output[key] = []
        if overlap:
            self.clock = T.max(t_issue, *busy.values())
        if when is not None and exact:
            for it in program.instructions:
                if it.done is None:
                    it.done = when
                elif it.done is not when:
                    it.done = ex._p_or(it.done, when)
        else:
            # a classical bit is dead once its predicated gates have been enqueued (stream order)
            for it in program.instructions:
                for b in (it.predicate, it.done):
                    if b is not None:
                        ex.creg.release(b.index)
        for b in temps:
            ex.creg.release(b.index)
        ex.programs_executed += 1
        ex._events += 1


class _Blocked(Exception):
    pass


class DQCExecutor:
    """Runs ``qpu_op_dict`` on ``backend`` following the reference's interpreter.

    Parameters
    ----------
    qpu_op_dict : dict   node name -> list of time slices -> list of op tuples
    dqc : DQCSpec
    backend : object implementing the engine vocabulary (``dqc_simulator_b200.engine.Engine``)
    forced_outcomes : optional iterable of 0/1/-1 consumed in measurement issue order
    host_branching : read outcomes on the host and branch like the reference (batch 1 only)
    """

    def __init__(self, qpu_op_dict, dqc, backend, forced_outcomes=None, host_branching=False,
                 swap_as_three_cnots=True, creg_bits=64, axis_pool=None, comm_axes_low=0,
                 parallel_programs=True, per_shot_time=True, predicated_time_weight=0.5, ancilla_free_cat=False):
        self.ops = qpu_op_dict
        self.dqc = dqc if isinstance(dqc, DQCSpec) else DQCSpec(list(dqc))
        self.backend = backend
        # ancilla_free_cat ("ancilla-free lowering"): on IDEAL hardware (no gate / measurement /
        # memory noise, pure Phi+ ebits) the cat-entangled comm qubit c' is an exact Z-basis copy of the data qubit
        # d, the gates it controls act as if d controlled them, and the disentangler (H, measure, Z) undoes the
        # copy -- so the whole primitive is "the remote gates, controlled by d": no ebit, no comm axis, no
        # measurement, no pass cut.  c' is then an ALIAS of d's engine axis.  The two cat measurement outcomes
        # are reported as 0 (they are fair coins that never reach the data state).  Refused on noisy hardware.
        self.ancilla_free_cat = bool(ancilla_free_cat)
        if self.ancilla_free_cat:
            self._require_ideal()
        self.aliased_gates = 0
        self.host_branching = host_branching
        self.swap_as_three_cnots = swap_as_three_cnots
        # True: instructions of one QuantumProgram on free positions overlap in simulated time (NetSquid's
        # parallel=True programs, see QPU.execute_program); False: strictly one after the other
        self.parallel_programs = parallel_programs
        # per_shot_time=True (default): every shot of a batch follows its own simulated timeline, exactly
        # like one reference trajectory (see the module docstring and timeline.py).
        # per_shot_time=False is the round-1 approximation kept for A/B measurements: the batch shares ONE
        # timeline that advances by weight x duration for a predicated correction (0.5 = the expected
        # timeline, cat / teleportation outcomes being 50:50) -- batch means are right to ~5e-6 but single
        # shots are off by ~1e-4 in fidelity.
        self.per_shot_time = bool(per_shot_time)
        self.predicated_time_weight = float(predicated_time_weight)
        self.T = tlm.Timeline(backend, None if self.per_shot_time else 0)
        self.creg = _CregAllocator(creg_bits)
        self._forced = list(forced_outcomes) if forced_outcomes is not None else None
        self._forced_i = 0
        self.qpus = {name: QPU(self.dqc.qpu(name), self) for name in qpu_op_dict}
        n_axes = backend.max_qubits if axis_pool is None else axis_pool
        self._n_axes = n_axes
        self._free_axes = list(range(n_axes))
        # comm_axes_low = k: the k lowest axes are reserved for communication qubits (and pinned in
        # the engine), so that the column AND row bits touched by every remote-gate primitive sit next
        # to the contiguous low block of the density matrix: longer HBM runs per tile
        self._comm_low = int(comm_axes_low)
        if self._comm_low:
            self._free_comm = list(range(self._comm_low))
            self._free_axes = list(range(self._comm_low, n_axes))
            if hasattr(backend, "set_option"):
                backend.set_option("pin_mask", (1 << self._comm_low) - 1)
        self.peak_axes = 0
        self.programs_executed = 0
        self.measurements = 0
        self.ebits = 0
        self.mailbox = {}     # (src, dst) -> list of (arrival_time, payload)
        self.ebit_requests = {}  # frozenset({a,b}) -> {node: (role, pos, time)}
        self.logged_results = []  # (node, position, CBit|int) in issue order
        self.logged_times = []    # simulated time at which each of them was produced
        self.heralded_links = []  # per heralded ebit: nodes, server, per-shot herald time, creg bit of the Bell state
        self.end_time = 0.0
        self.finished = False
        self._events = 0

    def _require_ideal(self):
        st = np.asarray(self.dqc.state4distribution, dtype=complex)
        ok = st.size == 4 and np.allclose(st.reshape(-1), [2 ** -0.5, 0, 0, 2 ** -0.5]) and self.dqc.heralded is None
        for sp in self.dqc.qpus:
            ok = ok and not (sp.p_depolar_error_cnot or sp.single_qubit_gate_error_prob or sp.meas_error_prob)
            ok = ok and not any(any(sp.mem_rates(p_)) for p_ in (0, sp.num_positions - 1))
        if not ok:
            raise ValueError("ancilla_free_cat needs ideal hardware: pure Phi+ ebits, no gate, measurement or memory "
                             "noise (on noisy hardware the comm qubits carry noise of their own)")

    # ---- resources -------------------------------------------------------------
    def _claim_axis(self, is_comm):
        """Processing qubits take the lowest free axis, comm qubits the highest: while a comm
        axis is free the active sub-block of the state stays contiguous in memory."""
        if self._comm_low and is_comm and self._free_comm:
            ax = self._free_comm.pop(0)
        else:
            if not self._free_axes:
                raise RuntimeError("engine register too small for the live qubits of this circuit")
            ax = self._free_axes.pop(-1 if is_comm else 0)
        used = self._n_axes - len(self._free_axes) - (len(self._free_comm) if self._comm_low else 0)
        self.peak_axes = max(self.peak_axes, used)
        return ax

    def _release_axis(self, ax):
        if self._comm_low and ax < self._comm_low:
            self._free_comm.append(ax)
            self._free_comm.sort()
            return
        self._free_axes.append(ax)
        self._free_axes.sort()

    def _next_forced(self):
        self.measurements += 1
        if self._forced is None:
            return -1
        if self._forced_i >= len(self._forced):
            return -1
        v = self._forced[self._forced_i]
        self._forced_i += 1
        return int(v)

    def _deliver(self, bit):
        if self.host_branching:
            v = int(self.backend.read_creg(bit)[0])
            self.creg.release(bit)
            return v
        return CBit(bit)

    # ---- derived classical bits (per-shot predicates) ------------------------------
    def _p_not(self, a, temps):
        b = CBit(self.creg.alloc())
        self.T.bit_op(tlm.CL_BNOT, b.index, a.index)
        temps.append(b)
        return b

    def _p_and(self, a, b, temps):
        """None = every shot."""
        if a is None or a is b:
            return b
        if b is None:
            return a
        c = CBit(self.creg.alloc())
        self.T.bit_op(tlm.CL_BAND, c.index, a.index, b.index)
        temps.append(c)
        return c

    def _p_or(self, a, b):
        c = CBit(self.creg.alloc())
        self.T.bit_op(tlm.CL_BOR, c.index, a.index, b.index)
        return c

    # ---- classical network ------------------------------------------------------
    def _send(self, src, dst, payload):
        t = self.T.add(self.qpus[src].clock, self.dqc.classical_delay)
        self.mailbox.setdefault((src, dst), []).append((t, payload))
        self._events += 1

    def _recv(self, src, dst):
        """generator: blocks until a message from src is available."""
        box = self.mailbox.setdefault((src, dst), [])
        while not box:
            yield "recv"
        t, payload = box.pop(0)
        self._events += 1
        q = self.qpus[dst]
        q.clock = self.T.max(q.clock, t)
        return payload

    def _request_entanglement(self, node, role, other, pos):
        """Rendezvous of the two link-layer requests, then the ebit lands on both comm
        positions (``dqc_control.py:198-264``; ``physical_layer.py:206-335, 451-479``).
        Timeline: client handshake -> server reply (one classical hop each way) ->
        source trigger -> both photons arrive ``ent_delay`` later."""
        key = frozenset((node, other))
        link = self.ebit_requests.setdefault(key, {"req": {}, "done": {}})
        link["req"][node] = (pos, self.qpus[node].clock, role)
        self._events += 1
        if other in link["req"]:
            first, second = sorted((node, other))   # qout0 -> side A, qout1 -> side B
            (pa, ta, ra), (pb, tb, rb) = link["req"][first], link["req"][second]
            qa, qb = self.qpus[first], self.qpus[second]
            qa._drop(pa)
            qb._drop(pb)
            ax_a = self._claim_axis(True)
            ax_b = self._claim_axis(True)
            link["req"] = {}
            self.ebits += 1
            if self.dqc.heralded is not None:
                server = second if rb == "server" or ra == "client" else first
                link["done"] = self._heralded_entangle(first, second, pa, pb, ax_a, ax_b, server,
                                                       self._handshake_done(ta, ra, tb, rb))
            else:
                t_arrive = self.T.add(self._handshake_done(ta, ra, tb, rb), self.dqc.ent_delay)
                self.backend.inject_2q(ax_a, ax_b, self.dqc.state4distribution, self.dqc.ebit_is_dm)
                qa.axis[pa], qb.axis[pb] = ax_a, ax_b
                qa.last_access[pa] = qb.last_access[pb] = t_arrive
                link["done"] = {first: t_arrive, second: t_arrive}
        while node not in link["done"]:
            yield "ebit"
        t_arrive = link["done"].pop(node)
        self._events += 1
        self.qpus[node].clock = self.T.max(self.qpus[node].clock, t_arrive)

    def _heralded_entangle(self, first, second, pa, pb, ax_a, ax_b, server, t_ready):
        """``MidpointHeraldingProtocol.entangle_comm_qubits`` (``physical_layer.py:685-753``) for every shot of the
        batch at once.  Per shot: the number of emit attempts until the midpoint BSM heralds success (geometric law,
        the shot's own Philox draw) moves that shot's clock by attempts x (INIT + EMIT + round trip); the heralded Bell
        state is Psi+ or Psi- with probability 1/2 each (a second draw); the server turns it into Phi+ with X, and Z
        where Psi- was heralded -- gates with their own noise and duration, in exactly those shots -- then tells the
        client (one classical hop).  Failed attempts leave nothing behind (the comm qubit is re-INITed); a shot that
        exhausts ``max_num_ent_attempts`` is treated as succeeding on the last one (the reference would signal
        ``ent_failed`` and stall; probability (1 - p)^100)."""
        be, T, hl = self.backend, self.T, self.dqc.heralded
        qa, qb = self.qpus[first], self.qpus[second]
        cycle = qa.spec.init_time + hl.emit_duration + hl.round_trip
        if not getattr(be, "n_time_regs", 0):
            raise NotImplementedError("the heralded link needs a backend with per-shot classical registers")
        t_herald, _k = T.attempts_then(t_ready, hl.success_probability, hl.max_num_ent_attempts, cycle)
        bit = self.creg.alloc()
        be.classical_rand(0, bit, 0.5, 0.0)                 # 1: Psi- heralded
        T.fence()
        self.creg.fence()
        psi_plus = np.array([0, 1, 1, 0], dtype=complex) / math.sqrt(2.0)
        be.inject_2q(ax_a, ax_b, psi_plus, False)
        be.set_predicate(bit)
        be.apply_1q(ax_b, np.diag([1.0 + 0j, -1.0]))        # Psi+ -> Psi- in the shots that heralded it (state preparation)
        be.set_predicate(-1)
        if hl.visibility < 1.0:
            if be.formalism not in (1, "dm"):
                raise NotImplementedError("heralded link: visibility < 1 needs the density-matrix formalism")
            be.pauli_channel(ax_b, [0.5 * (1.0 + hl.visibility), 0.0, 0.0, 0.5 * (1.0 - hl.visibility)])
        qa.axis[pa], qb.axis[pb] = ax_a, ax_b
        # the spin-photon pairs were emitted one round trip before the herald: memory noise runs from there
        qa.last_access[pa] = qb.last_access[pb] = T.add(t_herald, -hl.round_trip)
        qs, ps = (qa, pa) if server == first else (qb, pb)
        qs.clock = T.max(qs.clock, t_herald)
        fix = QuantumProgram()
        fix.apply(ins.INSTR_X, ps)
        fix.apply(ins.INSTR_Z, ps, predicate=CBit(bit))
        qs.execute_program(fix)
        t_server = qs.clock
        t_client = T.add(t_server, self.dqc.classical_delay)
        self.heralded_links.append(dict(nodes=(first, second), server=server, herald_time=t_herald, bell_bit=bit))
        return {first: t_server if server == first else t_client, second: t_server if server == second else t_client}

    def _handshake_done(self, ta, ra, tb, rb):
        """Time at which the client fires the source trigger (``physical_layer.py:269-335``): the
        client sends HANDSHAKE_READY and re-sends it every ``handshake_wait`` (600 us) until a copy
        reaches the server while the server is already listening; the server answers at once and
        the reply takes one more classical hop.  Trigger channel delay is 0
        (``connections.py:128-141``); both photons arrive ``ent_delay`` later."""
        d, w = self.dqc.classical_delay, self.dqc.handshake_wait
        if ra == "client" and rb != "client":
            tc, ts = ta, tb
        elif rb == "client" and ra != "client":
            tc, ts = tb, ta
        else:                       # symmetric request (not produced by the reference's compilers)
            return self.T.add(self.T.max(ta, tb), 2 * d)
        return self.T.handshake(tc, ts, d, w)

    # ---- interpreter ------------------------------------------------------------
    def _flush(self, q, program):
        q.execute_program(program)
        return QuantumProgram()

    def _apply_if(self, program, instruction, pos, outcome):
        """X / Z correction conditioned on a measurement outcome."""
        if isinstance(outcome, CBit):
            program.apply(instruction, pos, predicate=outcome)
        elif outcome == 1:
            program.apply(instruction, pos)
        elif outcome != 0:
            raise ValueError(
                "Measurement result must be an integer with value 0 or 1. The current value, "
                f"{outcome}, does not meet this criteria.")

    def _node_process(self, name, slice_ops):
        q = self.qpus[name]
        program = QuantumProgram()
        comm = float("nan")
        for op in slice_ops:
            last = op[-1]
            if len(op) == 2:                                   # single-qubit gate / INIT
                program.apply(_instr_of(op[0]), op[1], operator=_op_of(op[0]))
            elif last == "logged":                              # dqc_control.py:754-791
                program = self._flush(q, program)
                program.apply(_instr_of(op[0]), op[1], operator=_op_of(op[0]), output_key="result")
                q.execute_program(program)
                res = program.output["result"]
                self.logged_results.append((name, op[1], res[0] if res else None))
                self.logged_times.append(q.clock)
                program = QuantumProgram()
            elif last == "distribute_ebit":                     # dqc_control.py:803-832
                pos = q.comm_qubits_free[0]
                yield from self._request_entanglement(name, op[1], op[0], pos)
                del q.comm_qubits_free[0]
            elif isinstance(last, str) and op[-2] in ("cat", "tp"):
                if len(op) == 4:
                    data, other, scheme, role = op
                else:
                    data = None
                    other, scheme, role = op
                if scheme == "cat":
                    comm, program = yield from self._cat(q, role, program, other, data, comm)
                else:
                    comm, program = yield from self._tp(q, role, program, other, data, comm)
            elif isinstance(last, (int, np.integer)):           # local two-qubit gate
                q0 = comm if op[1] == -1 else op[1]
                program.apply(_instr_of(op[0]), [q0, op[2]], operator=_op_of(op[0]))
            else:
                raise ValueError(f"cannot interpret op {op!r}")
        self._flush(q, program)

    def _no_comm_free(self, q):
        if not q.comm_qubits_free:
            raise Exception(f"Scheduling error: No comm-qubits free. There are too many remote "
                            f"gates on {q.name} in this time slice")

    def _cat(self, q, role, program, other, data, comm):
        name = q.name
        if self.ancilla_free_cat:
            return (yield from self._cat_ancilla_free(q, role, program, other, data, comm))
        if role == "entangle":                                  # dqc_control.py:266-314
            program = self._flush(q, program)
            self._no_comm_free(q)
            c = q.comm_qubits_free[0]
            yield from self._request_entanglement(name, "client", other, c)
            program.apply(ins.INSTR_CNOT, [data, c])
            program.apply(ins.INSTR_MEASURE, [c], output_key="ma")
            q.execute_program(program)
            (ma,) = program.output["ma"]
            self._send(name, other, ma)
            return comm, QuantumProgram()
        if role == "correct":                                   # :316-359
            program = self._flush(q, program)
            self._no_comm_free(q)
            c = q.comm_qubits_free[0]
            yield from self._request_entanglement(name, "server", other, c)
            del q.comm_qubits_free[0]
            ma = yield from self._recv(other, name)
            self._apply_if(program, ins.INSTR_X, c, ma)
            return c, program
        if role == "disentangle_start":                         # :361-380
            program.apply(ins.INSTR_H, comm)
            program.apply(ins.INSTR_MEASURE, comm, output_key="mb")
            q.execute_program(program)
            (mb,) = program.output["mb"]
            self._send(name, other, mb)
            q.comm_qubits_free = sorted(q.comm_qubits_free + [comm])
            return comm, QuantumProgram()
        if role == "disentangle_end":                           # :382-403
            mb = yield from self._recv(other, name)
            if isinstance(mb, CBit) and self.per_shot_time:
                # the reference executes the pending program (local gates + this Z) only in the trajectories
                # with mb = 1; in the others the local gates stay pending until the next execute_program
                self._apply_if(program, ins.INSTR_Z, data, mb)
                q.execute_program(program, when=mb)
            elif isinstance(mb, CBit) or mb == 1:
                self._apply_if(program, ins.INSTR_Z, data, mb)
                q.execute_program(program)
                program = QuantumProgram()
            elif mb != 0:
                raise ValueError("Measurement result must be an integer with value 0 or 1. "
                                 f"The current value, {mb}, does not meet this criteria.")
            return comm, program
        raise ValueError("final element of gate_tuple is not valid role")

    def _cat_ancilla_free(self, q, role, program, other, data, comm):
        """The four cat-comm roles on ideal hardware, without ebit, comm axis or measurement (see __init__)."""
        name = q.name
        if role == "entangle":
            program = self._flush(q, program)
            self._no_comm_free(q)
            self._send(name, other, ("alias", q._materialise(data)))
            return comm, QuantumProgram()
        if role == "correct":
            program = self._flush(q, program)
            self._no_comm_free(q)
            c = q.comm_qubits_free[0]
            del q.comm_qubits_free[0]
            tag, ax = yield from self._recv(other, name)
            assert tag == "alias"
            q._drop(c)
            q.axis[c] = ax
            q.alias.add(c)
            q.last_access[c] = q.clock
            return c, program
        if role == "disentangle_start":
            program = self._flush(q, program)          # the gates the copy controls run before it is dropped
            q._drop(comm)
            self._send(name, other, 0)
            q.comm_qubits_free = sorted(q.comm_qubits_free + [comm])
            return comm, QuantumProgram()
        if role == "disentangle_end":
            mb = yield from self._recv(other, name)
            assert mb == 0
            return comm, program
        raise ValueError("final element of gate_tuple is not valid role")

    def _tp(self, q, role, program, other, data, comm):
        name = q.name
        if role == "bsm":                                       # dqc_control.py:405-489
            program = self._flush(q, program)
            self._no_comm_free(q)
            c = q.comm_qubits_free[0]
            if data == -1 and len(q.comm_qubits_free) < len(q.comm_qubit_positions):
                data = c - 1
            yield from self._request_entanglement(name, "client", other, c)
            program.apply(ins.INSTR_CNOT, [data, c])
            program.apply(ins.INSTR_H, [data])
            program.apply(ins.INSTR_MEASURE, [c], output_key="m1")
            program.apply(ins.INSTR_MEASURE, [data], output_key="m2")
            q.execute_program(program)
            (m1,) = program.output["m1"]
            (m2,) = program.output["m2"]
            program = QuantumProgram()
            program.apply(ins.INSTR_INIT, data)
            q.execute_program(program)
            self._send(name, other, (m1, m2))
            if data in q.comm_qubit_positions:
                q.comm_qubits_free = sorted(q.comm_qubits_free + [data])
            return comm, QuantumProgram()
        flags = {  # role -> (reserve_comm_qubit, swap_commNdata, manual)   :701-746
            "correct": (True, False, False),
            "correct4tele_only": (False, False, False),
            "swap_then_correct": (False, True, False),
            "correct_manually": (True, False, True),
        }
        if role not in flags:
            raise ValueError(f"final element, {role} of gate_tuple is not a valid role")
        reserve, swap, manual = flags[role]                     # :508-613
        program = self._flush(q, program)
        self._no_comm_free(q)
        c = data if manual else q.comm_qubits_free[0]
        if reserve:
            del q.comm_qubits_free[q.comm_qubits_free.index(c)]
        yield from self._request_entanglement(name, "server", other, c)
        m1, m2 = yield from self._recv(other, name)
        tgt = c
        if swap:
            program.apply(ins.INSTR_SWAP, [c, data])
            tgt = data
        self._apply_if(program, ins.INSTR_X, tgt, m1)
        self._apply_if(program, ins.INSTR_Z, tgt, m2)
        q.execute_program(program)
        return c, QuantumProgram()

    # ---- master ------------------------------------------------------------------
    def run(self):
        """Time-slice loop of ``DQCMasterProtocol.run`` (``dqc_control.py:1170-1202``): all
        QPUs start a slice together and the next slice starts when every QPU has finished."""
        n_slices = max(len(v) for v in self.ops.values())
        for s in range(n_slices):
            start = self.T.max(*[q.clock for q in self.qpus.values()])
            procs = {}
            for name, slices in self.ops.items():
                if s < len(slices):
                    self.qpus[name].clock = start
                    procs[name] = self._node_process(name, slices[s])
            while procs:
                progressed = False
                for name in list(procs):
                    try:
                        before = self._progress_marker()
                        next(procs[name])
                        if self._progress_marker() != before:
                            progressed = True
                    except StopIteration:
                        del procs[name]
                        progressed = True
                if not progressed and procs:
                    # one more full round with no state change anywhere = deadlock
                    if getattr(self, "_stalled", False):
                        raise UnfinishedQuantumCircuitError(
                            "Not all operations in the quantum circuit have run. Evaluation "
                            f"stopped in time slice {s} out of {n_slices - 1}; blocked QPUs: "
                            f"{sorted(procs)}")
                    self._stalled = True
                else:
                    self._stalled = False
        self.end_time = self.T.max(*[q.clock for q in self.qpus.values()])
        self.finished = True
        return self

    def _progress_marker(self):
        return self._events

    # ---- read-out ----------------------------------------------------------------
    def read_time(self, t):
        """A simulated time as one value per shot (numpy array of ``backend.batch``)."""
        if self.T.is_reg(t):
            return np.asarray(self.backend.read_time_reg(self.T.materialise(t).index), dtype=float)
        return np.full(getattr(self.backend, "batch", 1), float(t))

    def axes_of(self, qubits, apply_memory_noise=True):
        """[(position, node_name), ...] -> engine axes, like ``qmemory.pop(positions)`` at the
        end of the run (memory noise up to ``end_time``; ``util/helper.py:136-151``)."""
        axes = []
        for pos, node in qubits:
            q = self.qpus[node if isinstance(node, str) else node.name]
            if apply_memory_noise:
                (ax,) = q.touch([pos], self.end_time)
                q.last_access[pos] = self.end_time
            else:
                ax = q._materialise(pos)
            axes.append(ax)
        return axes

    def processing_qubits(self):
        """All initialised processing qubits, node by node (``run_experiment.py:38-43``)."""
        out = []
        for name, q in self.qpus.items():
            for pos in q.processing_qubit_positions:
                if q.occupied(pos):
                    out.append((pos, name))
        return out


def _instr_of(head):
    return head[0] if isinstance(head, (tuple, list)) else head


def _op_of(head):
    return head[1] if isinstance(head, (tuple, list)) else None
