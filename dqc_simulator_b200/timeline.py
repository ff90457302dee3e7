"""Per-shot simulated time for batched shots.

The reference runs ONE trajectory per ``ns.sim_run()``: a classically controlled X / Z correction
exists (135 us + its gate noise) only when the measurement outcome was 1
(``software/dqc_control.py:347-348, 387-390, 497-503``), ``_cat_disentangle_end`` executes the pending
program only in that case (``:382-403``), and the link layer re-sends its handshake on a 600 us grid
(``software/physical_layer.py:269-335``).  Every later memory-noise exponent
(``hardware/noise_models.py:248-253``: p = 1 - exp(-dt * 1e-9 * rate)) inherits those differences, so a
batch of shots has as many simulated timelines as it has outcome histories.

Here a simulated time is either a host ``float`` (identical for every shot of the batch: everything
before the first sampled measurement, and every forced / host-branching run) or a :class:`TReg`: a
per-shot double held in the backend's *classical register file* next to the per-shot creg word.  The
executor's clock arithmetic (``+``, ``max``, select on a classical bit, the handshake ``ceil``) goes
through :class:`Timeline`, which evaluates floats on the host and emits one classical instruction per
operation on registers.  The backend runs those instructions per shot (the CUDA engine inside its tile
passes, right after the measurement that produced the bits they read) and noise channels take their
probability from a register (``dyn_channel``).

Register discipline (the contract of ``dqe_classical_op``, include/dqc_engine.h): the engine may run the
classical instructions enqueued after a measurement EARLIER than the quantum ops enqueued between them
(it hoists them to just behind that measurement), so a register must not be overwritten while an op
enqueued earlier in the same span still reads it.  Freed registers therefore wait in ``_limbo`` until
the next measurement (``fence()``).
"""
import math

# classical opcodes (mirrored in csrc/dqe_device.cuh: ClOp)
(CL_CONST, CL_ADDI, CL_ADD, CL_SUB, CL_RSUBI, CL_MAX, CL_MAXI, CL_SEL, CL_CEILQ, CL_FMAI, CL_QEXP,
 CL_ADDIF) = range(1, 13)
CL_BAND, CL_BANDN, CL_BOR, CL_BNOT, CL_BCOPY = range(16, 21)

# dynamic one-qubit channels whose strength q comes from a register (csrc: DynKind)
DYN_DEPOL, DYN_DEPHASE, DYN_AMPDAMP, DYN_ZFLIP = 0, 1, 2, 3


class TReg:
    """Owner of classical register ``index`` (returned to the pool when the last reference dies)."""

    __slots__ = ("tl", "index")

    def __init__(self, tl, index):
        self.tl = tl
        self.index = index

    def __del__(self):
        tl = self.tl
        if tl is not None:
            tl._limbo.append(self.index)

    def __repr__(self):
        return f"TReg({self.index})"


class TVal:
    """Per-shot simulated time ``register + offset``.  Keeping the host-known offset out of the register
    makes most of the executor's clock arithmetic free: durations only move the offset, and two times
    that hang off the same register differ by a host constant (same idle time in every shot -> the
    memory-noise probability is a host constant again and the channel fuses like any other)."""

    __slots__ = ("reg", "off")

    def __init__(self, reg, off=0.0):
        self.reg = reg
        self.off = float(off)

    @property
    def index(self):
        return self.reg.index

    def __repr__(self):
        return f"TVal(r{self.reg.index}{self.off:+g})"


class Timeline:
    """Clock algebra over floats and :class:`TVal` values for one backend."""

    def __init__(self, backend, nregs=None):
        self.be = backend
        n = int(nregs if nregs is not None else getattr(backend, "n_time_regs", 0))
        self._free = list(range(n))
        self._limbo = []
        self.nregs = n
        self.peak = 0
        self.ops = 0

    # ---- registers ---------------------------------------------------------------------------
    def _new(self):
        if not self._free:
            if not self._limbo:
                raise RuntimeError(
                    "classical time registers exhausted: the circuit keeps more per-shot times alive than the "
                    f"backend's {self.nregs} registers" if self.nregs else
                    "this backend has no classical register file: batched shots with sampled mid-circuit "
                    "measurements need one (use host_branching=True, forced outcomes or per_shot_time=False)")
            # mid-span recycling: tell the backend not to hoist later classical ops over this point
            self.be.classical_fence()
            self.fence()
        r = TReg(self, self._free.pop(0))
        self.peak = max(self.peak, self.nregs - len(self._free))
        return r

    def fence(self):
        """A measurement (or an explicit classical fence) has been enqueued: registers freed before it
        may be reused by instructions enqueued after it."""
        if self._limbo:
            self._free.extend(self._limbo)
            self._limbo.clear()
            self._free.sort()

    def _emit(self, op, a=None, b=None, cbit=-1, imm=0.0, off=0.0):
        d = self._new()
        self.be.classical_op(op, d.index, a.index if a is not None else 0, b.index if b is not None else 0,
                             int(cbit), float(imm))
        self.ops += 1
        return TVal(d, off)

    def const(self, x):
        return self._emit(CL_CONST, imm=float(x))

    @staticmethod
    def is_reg(t):
        return isinstance(t, TVal)

    @staticmethod
    def same(a, b):
        """True when a and b are the same time in every shot (known on the host)."""
        if a is b:
            return True
        ra, rb = isinstance(a, TVal), isinstance(b, TVal)
        if ra and rb:
            return a.reg is b.reg and a.off == b.off
        return (not ra) and (not rb) and a == b

    def materialise(self, t):
        """A register holding exactly ``t`` (offset 0), e.g. for reading it back."""
        if isinstance(t, TVal):
            return t if t.off == 0.0 else self._emit(CL_ADDI, t, imm=t.off)
        return self.const(t)

    # ---- arithmetic ----------------------------------------------------------------------------
    def add(self, t, c):
        """t + c, c a host float."""
        if not isinstance(t, TVal):
            return t + c
        return t if c == 0 else TVal(t.reg, t.off + c)

    def sub(self, a, b):
        """a - b: a host float when both hang off the same register."""
        ra, rb = isinstance(a, TVal), isinstance(b, TVal)
        if not ra and not rb:
            return a - b
        if ra and rb:
            if a.reg is b.reg:
                return a.off - b.off
            return self._emit(CL_SUB, a, b, imm=a.off - b.off)
        if ra:
            return TVal(a.reg, a.off - b)
        return self._emit(CL_RSUBI, b, imm=a - b.off)

    def max(self, *ts):
        vals, c = [], None
        for t in ts:
            if isinstance(t, TVal):
                for i, v in enumerate(vals):
                    if v.reg is t.reg:
                        if t.off > v.off:
                            vals[i] = t
                        break
                else:
                    vals.append(t)
            else:
                c = t if c is None else max(c, t)
        if not vals:
            return c
        out = vals[0]
        for v in vals[1:]:   # max(Ra + oa, Rb + ob) = oa + max(Ra, Rb + (ob - oa))
            out = self._emit(CL_MAX, out, v, imm=v.off - out.off, off=out.off)
        if c is not None:
            out = self._emit(CL_MAXI, out, imm=c - out.off, off=out.off)
        return out

    def sel(self, pred, a, b):
        """``a`` in the shots whose classical bit ``pred`` is 1, ``b`` in the others; pred None = all."""
        if pred is None or self.same(a, b):
            return a
        ra, rb = isinstance(a, TVal), isinstance(b, TVal)
        if ra and rb and a.reg is b.reg:   # Rb + ob + (bit ? oa - ob : 0)
            return self._emit(CL_ADDIF, b, cbit=pred.index, imm=a.off - b.off, off=b.off)
        va = a if ra else self.const(a)
        vb = b if rb else self.const(b)
        return self._emit(CL_SEL, va, vb, cbit=pred.index, imm=va.off - vb.off, off=vb.off)

    def handshake(self, tc, ts, d, w):
        """Time at which the client fires the source trigger (``physical_layer.py:269-335``):
        tc + k w + 2 d with k = max(0, ceil((ts - tc - d) / w - 1e-12)) re-sends of HANDSHAKE_READY."""
        if not isinstance(tc, TVal) and not isinstance(ts, TVal):
            if w <= 0:
                return max(tc + d, ts) + d
            k = max(0, math.ceil((ts - tc - d) / w - 1e-12))
            return tc + k * w + 2 * d
        if w <= 0:
            return self.add(self.max(self.add(tc, d), ts), d)
        x = self.sub(ts, tc)
        if not isinstance(x, TVal):     # same register: the same number of re-sends in every shot
            k = max(0, math.ceil((x - d) / w - 1e-12))
            return self.add(tc, k * w + 2 * d)
        x = self.materialise(self.add(x, -d))
        k = self._emit(CL_CEILQ, x, imm=w)
        tcv = tc if isinstance(tc, TVal) else self.const(tc)
        return self._emit(CL_FMAI, k, tcv, imm=w, off=tcv.off + 2 * d)

    def qexp(self, dt, rate_per_ns):
        """q = 1 - exp(-max(0, dt) * rate) as a register (dt a per-shot value)."""
        return self._emit(CL_QEXP, self.materialise(dt), imm=rate_per_ns)

    def qexp_between(self, now, last, rate_per_ns):
        """q = 1 - exp(-max(0, now - last) * rate) as a register, in ONE instruction (at least one of the two
        times is per shot and they hang off different registers)."""
        ra, rb = isinstance(now, TVal), isinstance(last, TVal)
        off = (now.off if ra else now) - (last.off if rb else last)
        d = self._new()
        self.be.classical_qexp2(d.index, now.index if ra else 255, last.index if rb else 255, off, rate_per_ns)
        self.ops += 1
        return TVal(d, 0.0)

    def attempts_then(self, t0, p, kmax, cycle):
        """t0 + k * cycle with k = the shot's number of attempts until the first success of probability p (geometric
        law, capped at kmax; ``dqe_classical_rand`` kind 1): the heralded link's retry loop."""
        k = self._new()
        self.be.classical_rand(1, k.index, float(p), float(kmax))
        self.ops += 1
        base = t0 if isinstance(t0, TVal) else self.const(t0)
        return self._emit(CL_FMAI, TVal(k), base, imm=cycle, off=base.off), TVal(k)

    # ---- classical bits --------------------------------------------------------------------------
    def bit_op(self, op, dst_bit, a_bit, b_bit=0):
        self.be.classical_op(op, int(dst_bit), int(a_bit), int(b_bit), -1, 0.0)
        self.ops += 1
