"""ctypes binding of ``libdqcengine.so`` (C-ABI in ``include/dqc_engine.h``).

:class:`Engine` is the qstate backend the executor drives -- the stand-in for NetSquid's
``KetRepr`` / ``DenseDMRepr`` behind ``qubitapi``.  Method names and
argument meaning follow the op vocabulary of the boundary; every call lands in the CUDA
library.  There is **no CPU fallback**: if the shared library is missing or no CUDA device is
present, construction raises.

PyTorch is used only for plumbing: ``Engine.state_tensor()`` wraps the device buffer zero-copy
as a ``torch.complex128`` tensor, and an engine can adopt a caller-allocated torch tensor and
CUDA stream.
"""
import ctypes
import os

import numpy as np

KET = 0
DM = 1
# NetSquid's QFormalism members a user can select (docs/source/getting_started.rst:616-622 of the reference points at
# them): the two dense ones are the engine's own; SPARSEDM holds the same rho as DM (dense arithmetic here, identical
# results), STAB and GSLC hold stabiliser states, every one of which is a state vector -- Clifford circuits run on the
# ket engine unchanged (no tableau arithmetic is built: nothing dense to accelerate).
_FORMALISMS = {"ket": KET, "dm": DM, "sparsedm": DM, "stab": KET, "gslc": KET}


def resolve_formalism(f):
    """KET / DM constant for a formalism given as 0 / 1 or as a (NetSquid) name, case-insensitive."""
    if f in (KET, DM) and not isinstance(f, str):
        return int(f)
    name = str(getattr(f, "name", f)).lower()
    if name not in _FORMALISMS:
        raise ValueError(f"unknown state formalism {f!r}: one of {sorted(_FORMALISMS)}")
    return _FORMALISMS[name]

_HERE = os.path.dirname(os.path.abspath(__file__))
# DQE_LIB: another build of the same library (A/B measurements of kernel changes); never a fallback
LIB_PATH = os.environ.get("DQE_LIB") or os.path.join(_HERE, "lib", "libdqcengine.so")

_c_int_p = ctypes.POINTER(ctypes.c_int)
_c_dbl_p = ctypes.POINTER(ctypes.c_double)

# name -> (restype, argtypes); the complete list of symbols include/dqc_engine.h declares
ABI = {
    "dqe_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int,
                                  ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "dqe_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "dqe_reinit": (ctypes.c_int, [ctypes.c_void_p]),
    "dqe_last_error": (ctypes.c_char_p, [ctypes.c_void_p]),
    "dqe_version": (ctypes.c_char_p, []),
    "dqe_seed": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64]),
    "dqe_alloc_axis": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "dqe_free_axis": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "dqe_reset": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "dqe_live_mask": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64)]),
    "dqe_enqueue_1q": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _c_dbl_p]),
    "dqe_enqueue_2q": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, _c_dbl_p]),
    "dqe_enqueue_ctrl_1q": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, _c_dbl_p]),
    "dqe_enqueue_diag": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _c_int_p, _c_dbl_p]),
    "dqe_enqueue_kq": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _c_int_p, _c_dbl_p]),
    "dqe_peer_exchange": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int,
                                          ctypes.c_int, _c_int_p, _c_int_p]),
    "dqe_peer_stats": (ctypes.c_int, [ctypes.c_void_p, _c_dbl_p, ctypes.POINTER(ctypes.c_int64),
                                       ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]),
    "dqe_probs": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _c_dbl_p]),
    "dqe_probs_diag": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _c_int_p, _c_int_p, _c_int_p, ctypes.c_int, _c_dbl_p]),
    "dqe_force_live_mask": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint64]),
    "dqe_set_predicate": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "dqe_enqueue_cond_1q": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _c_dbl_p, ctypes.c_int]),
    "dqe_enqueue_pauli_channel": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _c_dbl_p]),
    "dqe_enqueue_pauli_channel_k": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _c_int_p, _c_dbl_p]),
    "dqe_enqueue_depolarize": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _c_int_p, ctypes.c_double]),
    "dqe_enqueue_kraus_1q": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, _c_dbl_p]),
    "dqe_inject_2q": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, _c_dbl_p, ctypes.c_int]),
    "dqe_inject_2q_mixture": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                             _c_dbl_p, _c_dbl_p]),
    "dqe_measure": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                   ctypes.c_int, ctypes.c_double]),
    "dqe_n_time_regs": (ctypes.c_int, []),
    "dqe_classical_op": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_double]),
    "dqe_classical_qexp2": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                           ctypes.c_double]),
    "dqe_classical_rand": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double]),
    "dqe_classical_fence": (ctypes.c_int, [ctypes.c_void_p]),
    "dqe_enqueue_dyn_channel": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "dqe_read_time_reg": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _c_dbl_p]),
    "dqe_read_outcomes": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                         ctypes.POINTER(ctypes.c_int8)]),
    "dqe_draw_count": (ctypes.c_int64, [ctypes.c_void_p]),
    "dqe_flush": (ctypes.c_int, [ctypes.c_void_p]),
    "dqe_replay": (ctypes.c_int, [ctypes.c_void_p]),
    "dqe_sync": (ctypes.c_int, [ctypes.c_void_p]),
    "dqe_read_creg": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int8)]),
    "dqe_read_creg_words": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64)]),
    "dqe_histogram": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _c_int_p, ctypes.POINTER(ctypes.c_int64)]),
    "dqe_reduced_dm": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _c_int_p, ctypes.c_int64, _c_dbl_p]),
    "dqe_fidelity_ket": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _c_int_p, _c_dbl_p, _c_dbl_p]),
    "dqe_get_state": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, _c_dbl_p]),
    "dqe_set_state": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, _c_dbl_p, ctypes.c_uint64]),
    "dqe_state_ptr": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p),
                                     ctypes.POINTER(ctypes.c_int64)]),
    "dqe_ipc_export": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "dqe_ipc_open": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]),
    "dqe_ipc_close": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "dqe_peer_1q": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, _c_dbl_p,
                                   ctypes.c_int]),
    "dqe_peer_swap": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "dqe_set_option": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int64]),
    "dqe_get_stats": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64)]),
    "dqe_reset_stats": (ctypes.c_int, [ctypes.c_void_p]),
    "dqe_time_passes": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "dqe_pass_time_ms": (ctypes.c_int, [ctypes.c_void_p, _c_dbl_p, ctypes.POINTER(ctypes.c_int64),
                                        ctypes.POINTER(ctypes.c_int64)]),
}

_lib = None


class EngineError(RuntimeError):
    pass


def load_library(path=None):
    """dlopen the engine and bind every symbol of the C-ABI.  Raises if the library was not
    built (``python -c 'import __graft_entry__ as g; g.build()'``) -- there is no fallback."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise EngineError(f"{p} not found: build the CUDA engine first (__graft_entry__.build())")
    lib = ctypes.CDLL(p)
    for name, (res, args) in ABI.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def _cbuf(a):
    """complex array -> contiguous interleaved doubles + pointer."""
    arr = np.ascontiguousarray(np.asarray(a, dtype=np.complex128))
    return arr, arr.ctypes.data_as(_c_dbl_p)


def _ibuf(a):
    arr = np.ascontiguousarray(np.asarray(list(a), dtype=np.int32))
    return arr, arr.ctypes.data_as(_c_int_p)


# C-ABI calls that only ENQUEUE (or launch) work and whose arguments do not depend on anything read back from the
# device: a run made of these alone can be recorded once and replayed on the same handle (Engine.record / play)
_REPLAYABLE = frozenset((
    "dqe_alloc_axis", "dqe_free_axis", "dqe_reset", "dqe_force_live_mask", "dqe_enqueue_1q", "dqe_enqueue_2q",
    "dqe_enqueue_ctrl_1q", "dqe_enqueue_diag", "dqe_enqueue_kq", "dqe_set_predicate", "dqe_enqueue_cond_1q",
    "dqe_enqueue_pauli_channel", "dqe_enqueue_pauli_channel_k", "dqe_enqueue_depolarize", "dqe_enqueue_kraus_1q", "dqe_classical_op",
    "dqe_classical_qexp2", "dqe_classical_fence", "dqe_classical_rand", "dqe_enqueue_dyn_channel", "dqe_inject_2q",
    "dqe_inject_2q_mixture", "dqe_measure", "dqe_flush", "dqe_sync", "dqe_peer_exchange"))
# queries that neither change the handle nor feed device data back into the program
_NEUTRAL = frozenset(("dqe_live_mask", "dqe_last_error", "dqe_get_stats", "dqe_draw_count", "dqe_n_time_regs",
                      "dqe_peer_stats", "dqe_pass_time_ms"))


class CallRecord:
    """The C-ABI call stream of one program on one handle (``Engine.record``): (function, argument tuple) pairs
    with their ctypes buffers kept alive.  ``valid`` turns False when the program called anything whose result
    could have steered it (a read-back): such a run cannot be replayed."""

    def __init__(self, owner=None):
        self.calls = []
        self.valid = True
        self.why = ""
        self.owner = owner      # handle value the calls are bound to

    def __len__(self):
        return len(self.calls)


class _RecordingLib:
    """Stands in for the CDLL while a program is recorded: every call goes through, enqueue-type calls are logged."""

    def __init__(self, lib, rec):
        self._lib, self._rec = lib, rec

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        if name in _NEUTRAL:
            return fn
        rec = self._rec
        if name not in _REPLAYABLE:
            def passthrough(*args):
                rec.valid = False
                rec.why = rec.why or name
                return fn(*args)
            return passthrough

        def logged(*args):
            rec.calls.append((fn, args))
            return fn(*args)
        return logged


class Engine:
    """Batched ket / density-matrix register on one B200.

    Same vocabulary as the test oracle (``oracle/numpy_engine.py``) so that the executor can
    drive either; here every method is a thin call through the C-ABI.
    """

    STAT_NAMES = ("passes", "micro_ops", "kernel_launches", "amp_updates", "api_ops", "finish_launches",
                  "h2d_bytes", "merged_ops")

    def __init__(self, formalism, max_qubits, batch=1, seed=0, shot_offset=0, device=0,
                 state_tensor=None, stream=None, tile_bits=None, fuse=True):
        self._h = ctypes.c_void_p()
        self._lib = load_library()
        self.formalism = resolve_formalism(formalism)
        self.max_qubits = int(max_qubits)
        self.batch = int(batch)
        self.device = int(device)
        self._keepalive = state_tensor
        if state_tensor is not None:
            # a caller-provided buffer is written by memset / kernels: a short, strided, host or wrong-device
            # tensor would mean out-of-bounds device writes
            import torch

            need = self.batch * (2 ** self.max_qubits if self.formalism == KET else 4 ** self.max_qubits)
            if not (isinstance(state_tensor, torch.Tensor) and state_tensor.is_cuda):
                raise EngineError("state_tensor must be a CUDA torch tensor")
            if state_tensor.device.index != self.device:
                raise EngineError(f"state_tensor lives on cuda:{state_tensor.device.index}, the engine on cuda:{self.device}")
            if state_tensor.dtype != torch.complex128 or not state_tensor.is_contiguous():
                raise EngineError("state_tensor must be a contiguous complex128 tensor")
            if state_tensor.numel() < need:
                raise EngineError(f"state_tensor holds {state_tensor.numel()} amplitudes, the register needs {need}")
        ext = ctypes.c_void_p(state_tensor.data_ptr()) if state_tensor is not None else None
        st = ctypes.c_void_p(int(stream)) if stream else None
        rc = self._lib.dqe_create(ctypes.byref(self._h), self.formalism, self.max_qubits, self.batch,
                                  self.device, ext, st)
        if rc != 0:
            msg = self._lib.dqe_last_error(None).decode()
            self._h = ctypes.c_void_p()
            raise EngineError(f"dqe_create failed ({rc}): {msg}")
        self._ck(self._lib.dqe_seed(self._h, int(seed) & 0xFFFFFFFFFFFFFFFF, int(shot_offset)))
        self.n_time_regs = int(self._lib.dqe_n_time_regs())
        if tile_bits is None and os.environ.get("DQE_TILE_BITS"):
            tile_bits = int(os.environ["DQE_TILE_BITS"])
        if tile_bits is not None:
            self.set_option("tile_bits", tile_bits)
        if not fuse or os.environ.get("DQE_FUSE") == "0":
            self.set_option("fuse", 0)
        for key in ("merge", "group", "tma", "min_run_bits", "threads", "ctas_per_sm", "persistent", "cluster"):  # A/B switches for profiling
            v = os.environ.get("DQE_" + key.upper())
            if v is not None:
                self.set_option(key, int(v))

    # ------------------------------------------------------------------ plumbing
    def _ck(self, rc):
        if rc != 0:
            raise EngineError(f"engine error {rc}: {self._lib.dqe_last_error(self._h).decode()}")

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.dqe_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ recorded programs
    def record(self):
        """Context manager: log the C-ABI call stream of the program run inside it (executor walk, adapter logic,
        argument marshalling all happen once).  ``play`` then re-issues the stream on this handle after a ``reinit``
        -- the host-side plan cache for programs that span several flushes (sharded state vectors: passes, exchange,
        passes, ...); ``dqe_replay`` is the in-library one for a single flush."""
        import contextlib

        @contextlib.contextmanager
        def cm():
            rec = CallRecord(self._h.value)
            real = self._lib
            self._lib = _RecordingLib(real, rec)
            try:
                yield rec
            finally:
                self._lib = real
        return cm()

    def play(self, rec):
        """Re-issue a recorded call stream (same handle, same initial register state as when it was recorded)."""
        if not rec.valid:
            raise EngineError(f"this run read device data back ({rec.why}): it cannot be replayed")
        if not self._h.value or rec.owner != self._h.value:
            raise EngineError("a recorded program replays on the (open) handle it was recorded on only")
        for fn, args in rec.calls:
            if fn(*args) != 0:
                raise EngineError(f"engine error during replay: {self._lib.dqe_last_error(self._h).decode()}")

    def reinit(self):
        """Fresh |0..0> batch on the same buffers (next batch of shots)."""
        self._ck(self._lib.dqe_reinit(self._h))

    def set_option(self, key, value):
        self._ck(self._lib.dqe_set_option(self._h, key.encode(), int(value)))
        if key != "plan_cache":
            self._shot_batch_plan = None   # the library dropped its cached plan too (experiment.run_shot_batch)

    def seed(self, seed, shot_offset=0):
        self._ck(self._lib.dqe_seed(self._h, int(seed) & 0xFFFFFFFFFFFFFFFF, int(shot_offset)))

    @property
    def live(self):
        m = ctypes.c_uint64()
        self._ck(self._lib.dqe_live_mask(self._h, ctypes.byref(m)))
        return [a for a in range(self.max_qubits) if (m.value >> a) & 1]

    def force_live(self, axes):
        """Declare the live axes after a data movement outside the gate queue (every other axis must be in |0>)."""
        mask = 0
        for a in axes:
            mask |= 1 << int(a)
        self._ck(self._lib.dqe_force_live_mask(self._h, ctypes.c_uint64(mask)))

    @property
    def n_live(self):
        return len(self.live)

    # ------------------------------------------------------------------ register
    def alloc_axis(self, axis=-1):
        if axis < 0:
            live = set(self.live)
            axis = min(a for a in range(self.max_qubits) if a not in live)
        self._ck(self._lib.dqe_alloc_axis(self._h, int(axis)))
        return axis

    def free_axis(self, axis):
        self._ck(self._lib.dqe_free_axis(self._h, int(axis)))

    def reset(self, axis):
        self._ck(self._lib.dqe_reset(self._h, int(axis)))

    # ------------------------------------------------------------------ unitaries
    def apply_1q(self, axis, m):
        _k, p = _cbuf(np.asarray(m, dtype=complex).reshape(2, 2))
        self._ck(self._lib.dqe_enqueue_1q(self._h, int(axis), p))

    def apply_2q(self, a, b, m):
        _k, p = _cbuf(np.asarray(m, dtype=complex).reshape(4, 4))
        self._ck(self._lib.dqe_enqueue_2q(self._h, int(a), int(b), p))

    def apply_ctrl_1q(self, c, t, m):
        _k, p = _cbuf(np.asarray(m, dtype=complex).reshape(2, 2))
        self._ck(self._lib.dqe_enqueue_ctrl_1q(self._h, int(c), int(t), p))

    def apply_diag(self, axes, d):
        _a, pa = _ibuf(axes)
        _k, p = _cbuf(np.asarray(d, dtype=complex).reshape(-1))
        self._ck(self._lib.dqe_enqueue_diag(self._h, len(_a), pa, p))

    def apply_kq(self, axes, m):
        """Dense k-qubit gate (k <= 4, first axis = most significant index): FP64 tensor cores for k >= 3."""
        _a, pa = _ibuf(axes)
        k = len(_a)
        _k, p = _cbuf(np.asarray(m, dtype=complex).reshape(2 ** k, 2 ** k))
        self._ck(self._lib.dqe_enqueue_kq(self._h, k, pa, p))

    def set_predicate(self, creg_bit):
        self._ck(self._lib.dqe_set_predicate(self._h, int(creg_bit)))

    def cond_1q(self, axis, m, creg_bit):
        _k, p = _cbuf(np.asarray(m, dtype=complex).reshape(2, 2))
        self._ck(self._lib.dqe_enqueue_cond_1q(self._h, int(axis), p, int(creg_bit)))

    # ------------------------------------------------------------------ channels
    def pauli_channel(self, axis, p):
        """One axis and (pI, pX, pY, pZ), or a list of two axes and 16 weights p[4 * (Pauli on axes[0]) + (Pauli on
        axes[1])] -- any correlated two-qubit Pauli noise (``dqe_enqueue_pauli_channel_k``)."""
        if isinstance(axis, (list, tuple, np.ndarray)):
            _a, pa = _ibuf(axis)
            arr = np.ascontiguousarray(np.asarray(p, dtype=np.float64).reshape(4 ** len(_a)))
            self._ck(self._lib.dqe_enqueue_pauli_channel_k(self._h, len(_a), pa, arr.ctypes.data_as(_c_dbl_p)))
            return
        arr = np.ascontiguousarray(np.asarray(p, dtype=np.float64).reshape(4))
        self._ck(self._lib.dqe_enqueue_pauli_channel(self._h, int(axis), arr.ctypes.data_as(_c_dbl_p)))

    def depolarize(self, axes, p):
        _a, pa = _ibuf(axes)
        self._ck(self._lib.dqe_enqueue_depolarize(self._h, len(_a), pa, float(p)))

    def kraus_1q(self, axis, ks):
        arr = np.asarray(ks, dtype=complex).reshape(-1, 2, 2)
        _k, p = _cbuf(arr)
        self._ck(self._lib.dqe_enqueue_kraus_1q(self._h, int(axis), arr.shape[0], p))

    # ------------------------------------------------------------------ per-shot classical machine
    def classical_op(self, opcode, dst, a=0, b=0, cbit=-1, imm=0.0):
        """One instruction of the per-shot register machine (opcodes: ``timeline.CL_*``)."""
        self._ck(self._lib.dqe_classical_op(self._h, int(opcode), int(dst), int(a), int(b), int(cbit), float(imm)))

    def classical_qexp2(self, dst, a, b, offset, rate):
        """dst <- 1 - exp(-max(0, R[a] - R[b] + offset) * rate); a / b = 255 stand for 0."""
        self._ck(self._lib.dqe_classical_qexp2(self._h, int(dst), int(a), int(b), float(offset), float(rate)))

    def classical_fence(self):
        self._ck(self._lib.dqe_classical_fence(self._h))

    def classical_rand(self, kind, dst, p, kmax=0.0):
        """kind 0: creg bit dst <- (u < p); kind 1: register dst <- attempts until the first success of probability p
        (capped at kmax); u = the shot's next Philox uniform."""
        self._ck(self._lib.dqe_classical_rand(self._h, int(kind), int(dst), float(p), float(kmax)))

    def dyn_channel(self, axis, kind, reg):
        """One-qubit channel whose strength is read per shot from classical register ``reg``
        (``timeline.DYN_DEPOL / DYN_DEPHASE / DYN_AMPDAMP``)."""
        self._ck(self._lib.dqe_enqueue_dyn_channel(self._h, int(axis), int(kind), int(reg)))

    def read_time_reg(self, reg):
        out = np.empty(self.batch, dtype=np.float64)
        self._ck(self._lib.dqe_read_time_reg(self._h, int(reg), out.ctypes.data_as(_c_dbl_p)))
        return out

    def read_outcomes(self):
        """(number of measurements so far, batch) int8 -- needs ``set_option("outcome_log", n)``."""
        n = int(self._lib.dqe_draw_count(self._h))
        out = np.empty((n, self.batch), dtype=np.int8)
        if n:
            self._ck(self._lib.dqe_read_outcomes(self._h, 0, n, out.ctypes.data_as(ctypes.POINTER(ctypes.c_int8))))
        return out[(out >= 0).all(axis=1)]   # draws that were not measurements (sampled noise) hold -1

    # ------------------------------------------------------------------ injection
    def inject_2q(self, a, b, state, is_dm):
        """``state``: ket (4,) or density matrix (4, 4); a = most significant qubit."""
        if self.formalism == KET and is_dm:
            # spectral decomposition on the host, largest eigenvalue first, so that the
            # per-shot component draw is reproducible
            sig = np.asarray(state, dtype=complex).reshape(4, 4)
            lam, vec = np.linalg.eigh(sig)
            lam = np.clip(lam[::-1], 0, None)
            vec = vec[:, ::-1]
            probs = np.ascontiguousarray(lam / lam.sum(), dtype=np.float64)
            _k, pk = _cbuf(vec.T)
            self._ck(self._lib.dqe_inject_2q_mixture(self._h, int(a), int(b), 4,
                                                     probs.ctypes.data_as(_c_dbl_p), pk))
            return
        arr = np.asarray(state, dtype=complex).reshape(16 if is_dm else 4)
        _k, p = _cbuf(arr)
        self._ck(self._lib.dqe_inject_2q(self._h, int(a), int(b), p, 1 if is_dm else 0))

    # ------------------------------------------------------------------ measurement
    def measure(self, axis, creg_bit, forced=-1, discard=False, flip_prob=0.0):
        cb = -1 if creg_bit is None else int(creg_bit)
        f = -1 if forced is None else int(forced)
        self._ck(self._lib.dqe_measure(self._h, int(axis), cb, f, 1 if discard else 0, float(flip_prob)))

    # ------------------------------------------------------------------ execution / read-out
    def flush(self):
        self._ck(self._lib.dqe_flush(self._h))

    def sync(self):
        self._ck(self._lib.dqe_sync(self._h))

    def replay(self):
        """Launch the cached plan of the last whole-program flush on the freshly reinitialised register
        (option ``plan_cache``; include/dqc_engine.h: dqe_replay)."""
        self._ck(self._lib.dqe_replay(self._h))

    def read_creg(self, bit):
        out = np.empty(self.batch, dtype=np.int8)
        self._ck(self._lib.dqe_read_creg(self._h, int(bit), out.ctypes.data_as(ctypes.POINTER(ctypes.c_int8))))
        return out

    def read_creg_words(self):
        out = np.empty(self.batch, dtype=np.uint64)
        self._ck(self._lib.dqe_read_creg_words(self._h, out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))))
        return out

    def histogram(self, cbits):
        _a, pa = _ibuf(cbits)
        out = np.zeros(2 ** len(_a), dtype=np.int64)
        self._ck(self._lib.dqe_histogram(self._h, len(_a), pa, out.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))))
        return out

    def reduced_dm(self, axes, shot=0):
        _a, pa = _ibuf(axes)
        k = len(_a)
        out = np.empty((2 ** k, 2 ** k), dtype=np.complex128)
        self._ck(self._lib.dqe_reduced_dm(self._h, k, pa, int(shot), out.ctypes.data_as(_c_dbl_p)))
        return out

    def fidelity_ket_all(self, axes, ket):
        _a, pa = _ibuf(axes)
        _k, pk = _cbuf(np.asarray(ket, dtype=complex).reshape(-1))
        out = np.empty(self.batch, dtype=np.float64)
        self._ck(self._lib.dqe_fidelity_ket(self._h, len(_a), pa, pk, out.ctypes.data_as(_c_dbl_p)))
        return out

    def fidelity_ket(self, axes, ket, shot=0):
        return float(self.fidelity_ket_all(axes, ket)[shot])

    def full_state(self, shot=None):
        """Raw fixed-layout state: (b, 2^n) ket or (b, 2^n, 2^n) dm, like the oracle's."""
        shots = range(self.batch) if shot is None else [shot]
        n = self.max_qubits
        dim = 2 ** n if self.formalism == KET else 4 ** n
        out = np.empty((len(shots), dim), dtype=np.complex128)
        for i, s in enumerate(shots):
            self._ck(self._lib.dqe_get_state(self._h, int(s), out[i].ctypes.data_as(_c_dbl_p)))
        if self.formalism == DM:
            out = out.reshape(len(shots), 2 ** n, 2 ** n)
        return out

    def set_state(self, shot, raw, live_axes):
        arr, p = _cbuf(np.asarray(raw, dtype=complex).reshape(-1))
        mask = 0
        for a in live_axes:
            mask |= 1 << int(a)
        self._ck(self._lib.dqe_set_state(self._h, int(shot), p, mask))

    def state_tensor(self):
        """Zero-copy torch view of the device buffer (complex128, flat)."""
        import torch

        dev = ctypes.c_void_p()
        nbytes = ctypes.c_int64()
        self._ck(self._lib.dqe_state_ptr(self._h, ctypes.byref(dev), ctypes.byref(nbytes)))
        n = nbytes.value // 16

        class _Holder:
            __cuda_array_interface__ = {
                "shape": (n * 2,), "typestr": "<f8", "data": (dev.value, False), "version": 2}

        t = torch.as_tensor(_Holder(), device=f"cuda:{self.device}")
        return torch.view_as_complex(t.view(n, 2))

    # ------------------------------------------------------------------ peer memory (sharded kets)
    def ipc_export(self):
        buf = ctypes.create_string_buffer(64)
        self._ck(self._lib.dqe_ipc_export(self._h, buf))
        return buf.raw

    def ipc_open(self, handle):
        p = ctypes.c_void_p()
        self._ck(self._lib.dqe_ipc_open(self._h, ctypes.c_char_p(handle), ctypes.byref(p)))
        return p.value

    def ipc_close(self, ptr):
        self._ck(self._lib.dqe_ipc_close(self._h, ctypes.c_void_p(ptr)))

    def peer_1q(self, peer_ptr, my_bit, half, m, ctrl_axis=-1):
        _k, p = _cbuf(np.asarray(m, dtype=complex).reshape(2, 2))
        self._ck(self._lib.dqe_peer_1q(self._h, ctypes.c_void_p(peer_ptr), int(my_bit), int(half), p, int(ctrl_axis)))

    def peer_swap(self, peer_ptr, my_bit, half, local_axis):
        self._ck(self._lib.dqe_peer_swap(self._h, ctypes.c_void_p(peer_ptr), int(my_bit), int(half), int(local_axis)))

    def peer_exchange(self, peer_ptrs, rank, rank_bits, local_axes):
        """Multi-axis global <-> local exchange with every partner in one launch (k_peer_exchange); enqueue only."""
        world = len(peer_ptrs)
        arr = (ctypes.c_void_p * world)(*[ctypes.c_void_p(p) if p else None for p in peer_ptrs])
        _a, pj = _ibuf(rank_bits)
        _b, pl = _ibuf(local_axes)
        self._ck(self._lib.dqe_peer_exchange(self._h, arr, int(rank), world, len(_a), pj, pl))

    def peer_stats(self):
        ms, cnt, nbytes, err = ctypes.c_double(), ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        self._ck(self._lib.dqe_peer_stats(self._h, ctypes.byref(ms), ctypes.byref(cnt), ctypes.byref(nbytes), ctypes.byref(err)))
        return dict(ms=ms.value, count=cnt.value, bytes_sent=nbytes.value, error=err.value)

    def probs(self, axis):
        out = (ctypes.c_double * 2)()
        self._ck(self._lib.dqe_probs(self._h, int(axis), out))
        return float(out[0]), float(out[1])

    def probs_diag(self, pairs, target):
        """Density matrix as a ket on (rows | columns): sum of this shard's diagonal elements by the value of local bit
        ``target``; pairs = [(a, b, fixed)], fixed < 0: local bits a and b must agree, else bit a must equal fixed."""
        _a, pa = _ibuf([p[0] for p in pairs] or [0])
        _b, pb = _ibuf([p[1] for p in pairs] or [0])
        _f, pf = _ibuf([p[2] for p in pairs] or [0])
        out = (ctypes.c_double * 2)()
        self._ck(self._lib.dqe_probs_diag(self._h, len(pairs), pa, pb, pf, int(target), out))
        return float(out[0]), float(out[1])

    # ------------------------------------------------------------------ instrumentation
    def stats(self):
        out = (ctypes.c_int64 * 8)()
        self._ck(self._lib.dqe_get_stats(self._h, out))
        return dict(zip(self.STAT_NAMES, list(out)[:8]))

    def reset_stats(self):
        self._ck(self._lib.dqe_reset_stats(self._h))

    def time_passes(self, enable=True):
        self._ck(self._lib.dqe_time_passes(self._h, 1 if enable else 0))

    def pass_time(self):
        ms = ctypes.c_double()
        n = ctypes.c_int64()
        amps = ctypes.c_int64()
        self._ck(self._lib.dqe_pass_time_ms(self._h, ctypes.byref(ms), ctypes.byref(n), ctypes.byref(amps)))
        return ms.value, n.value, amps.value
