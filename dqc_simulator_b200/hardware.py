"""Hardware specification consumed by the executor: which noise channel and which simulated
duration attach to which instruction.

Mirrors the *parameters* of the reference's hardware layer without any of its DES plumbing:

* :class:`QPUSpec`  <- ``QPU`` / ``NoisyQPU`` (``hardware/quantum_processors.py:83-123,
  796-909``): memory layout (comm positions first, then processing,
  ``quantum_processors.py:147-181``), the physical-instruction table of
  ``_phys_instructions_4_standard_lib_gates_and_convenience_ops`` (``:306-457``) and the
  SWAP = 3 CNOT composite (``:901-909``).
* :class:`DQCSpec`  <- ``DQC`` (``hardware/dqc_creation.py:443-617``) with a
  ``BlackBoxEntanglingQsourceConnection`` (``hardware/connections.py:85-172``):
  ``state4distribution`` + ``delay`` and 2 m classical fibres (``dqc_creation.py:512-516``;
  ``FibreDelayModel`` default c = 200 000 km/s).
"""
from dataclasses import dataclass, field

import numpy as np

_S2 = 1.0 / np.sqrt(2.0)
B00 = np.array([_S2, 0, 0, _S2], dtype=complex)  # ks.b00


def werner_state(F):
    """``qlib/states.py:63-84``: F |b00><b00| + (1-F)/3 (|b01><b01| + |b10><b10| + |b11><b11|)."""
    b01 = np.array([0, _S2, _S2, 0], dtype=complex)
    b10 = np.array([_S2, 0, 0, -_S2], dtype=complex)
    b11 = np.array([0, _S2, -_S2, 0], dtype=complex)
    out = F * np.outer(B00, B00.conj())
    for b in (b01, b10, b11):
        out = out + (1 - F) / 3 * np.outer(b, b.conj())
    return out


# ---- memory noise models a user can hand to a QPU (``mem_noise_models`` / ``comm_qubit_mem_noise_model`` /
# ``processing_qubit_mem_noise_model``, quantum_processors.py:83-123, 183-221, 257-260).  Same
# constructor arguments as NetSquid's ``netsquid.components.models.qerrormodels`` classes [NS-doc]; here they are
# parameter holders that the executor turns into idle-time channels of the engine.
@dataclass
class DepolarNoiseModel:
    """(1 - p) rho + p I/2 with p = 1 - exp(-dt * 1e-9 * depolar_rate)  (rate in Hz), or p = depolar_rate when
    ``time_independent``."""
    depolar_rate: float = 0.0
    time_independent: bool = False


@dataclass
class DephaseNoiseModel:
    """Z with probability p = 1 - exp(-dt * 1e-9 * dephase_rate)  (``qapi.dephase(qubit, prob)``), or
    p = dephase_rate when ``time_independent``."""
    dephase_rate: float = 0.0
    time_independent: bool = False


@dataclass
class T1T2NoiseModel:
    """Amplitude damping with gamma = 1 - exp(-dt / T1), then pure dephasing that brings the total coherence decay
    to exp(-dt / T2): off-diagonals x exp(-dt (1/T2 - 1/(2 T1))).  T1, T2 in ns (0 = off), T2 <= 2 T1."""
    T1: float = 0.0
    T2: float = 0.0


def mem_channel_rates(model):
    """Noise model -> idle-time channel rates in Hz: dict(depol, dephase (coherence decay), ampdamp, zflip)."""
    out = dict(depol=0.0, dephase=0.0, ampdamp=0.0, zflip=0.0)
    if model is None:
        return out
    name = type(model).__name__
    if name == "DepolarNoiseModel":
        if getattr(model, "time_independent", False):
            raise ValueError("time-independent memory noise has no idle-time meaning")
        out["depol"] = float(model.depolar_rate)
    elif name == "DephaseNoiseModel":
        if getattr(model, "time_independent", False):
            raise ValueError("time-independent memory noise has no idle-time meaning")
        out["zflip"] = float(model.dephase_rate)
    elif name == "T1T2NoiseModel":
        T1, T2 = float(model.T1), float(model.T2)
        if T1 > 0:
            out["ampdamp"] = 1e9 / T1
        if T2 > 0:
            r = 1e9 / T2 - (0.5e9 / T1 if T1 > 0 else 0.0)
            if r < -1e-12:
                raise ValueError("T1T2NoiseModel: T2 must not exceed 2 * T1")
            out["dephase"] = max(r, 0.0)
    else:
        raise TypeError(f"unsupported memory noise model {name}")
    return out


@dataclass
class PhysInstr:
    duration: float
    noise: str = "none"  # "none" | "depol1" | "depolk" | "flip_before"
    prob: float = 0.0
    discard: bool = False


@dataclass
class QPUSpec:
    """Defaults are ``NoisyQPU``'s (``quantum_processors.py:849-866``)."""
    name: str = "node_0"
    num_positions: int = 20
    num_comm_qubits: int = 2
    p_depolar_error_cnot: float = 0.0
    single_qubit_gate_error_prob: float = 0.0
    meas_error_prob: float = 0.0
    comm_qubit_depolar_rate: float = 0.0
    proc_qubit_depolar_rate: float = 0.0
    single_qubit_gate_time: float = 135e3
    two_qubit_gate_time: float = 600e3
    measurement_time: float = 600e4
    init_time: float = 3.0
    measure_discards: bool = True
    # optional extra memory channel per position type (north_star: dephasing / amplitude
    # damping).  Not wired by any reference hardware spec; rates in Hz.
    mem_dephase_rate: float = 0.0
    mem_amp_damp_rate: float = 0.0
    # user-selected models per position type / per position, like the reference's QPU arguments
    # (quantum_processors.py:83-123): DepolarNoiseModel, DephaseNoiseModel or T1T2NoiseModel (above, or NetSquid's
    # own objects: duck-typed on the class name and attributes).  They ADD to the rate fields.
    comm_qubit_mem_noise_model: object = None
    processing_qubit_mem_noise_model: object = None
    mem_noise_models: list = None

    @property
    def comm_qubit_positions(self):
        return list(range(self.num_comm_qubits))

    @property
    def processing_qubit_positions(self):
        return list(range(self.num_comm_qubits, self.num_positions))

    def mem_rate(self, pos):
        return self.comm_qubit_depolar_rate if pos < self.num_comm_qubits else self.proc_qubit_depolar_rate

    def mem_rates(self, pos):
        """Idle-time channel rates (Hz) of memory position ``pos``: (depol, dephase, ampdamp, zflip)."""
        if self.mem_noise_models is not None:
            model = self.mem_noise_models[pos]
        else:
            model = self.comm_qubit_mem_noise_model if pos < self.num_comm_qubits else self.processing_qubit_mem_noise_model
        m = mem_channel_rates(model)
        return (self.mem_rate(pos) + m["depol"], self.mem_dephase_rate + m["dephase"],
                self.mem_amp_damp_rate + m["ampdamp"], m["zflip"])

    def phys(self, token):
        """Row of the physical-instruction table for an instruction token."""
        if token.kind == "init":
            return PhysInstr(self.init_time)
        if token.kind == "measure":
            return PhysInstr(self.measurement_time, "flip_before", self.meas_error_prob,
                             discard=self.measure_discards)
        if token.kind == "discard":
            return PhysInstr(self.single_qubit_gate_time, discard=True)
        # INSTR_SINGLE_QUBIT_NEGLIGIBLE_TIME / INSTR_TWO_QUBIT_NEGLIGIBLE_TIME (qlib/gates.py:112-115, spelt
        # "neglibible_time_instr" / "neglible_time_2_qubit" upstream): duration 3 ns, no noise model
        # (quantum_processors.py:450-455).  Matched by the token names the reference uses.
        if token.name in ("neglibible_time_instr", "neglible_time_2_qubit"):
            return PhysInstr(3.0)
        if token.num_qubits == 1:
            return PhysInstr(self.single_qubit_gate_time, "depol1", self.single_qubit_gate_error_prob)
        return PhysInstr(self.two_qubit_gate_time, "depolk", self.p_depolar_error_cnot)


@dataclass
class HeraldedLinkSpec:
    """Probabilistic entangling link: ``create_midpoint_heralded_entangling_link`` (``hardware/connections.py:399-499``,
    a ``netsquid_physlayer`` ``MiddleHeraldedConnection``) driven by ``MidpointHeraldingProtocol``
    (``software/physical_layer.py:596-768``).  Each attempt: INIT the comm qubit, ``INSTR_EMIT`` a photon entangled
    with it on both QPUs, the photons meet at the midpoint Bell-state measurement, the herald comes back.  A linear-
    optics BSM recognises |Psi+> and |Psi-> only, so an attempt succeeds with probability 1/2 times the two arms'
    transmission; the server then turns the heralded state into |Phi+> with X (and Z for Psi-).
    Model [NS-doc, pinned by the reference only in the ideal case, test_physical_layer_and_hardware_integration.py:
    267-295]: double-click protocol without dark counts; photon loss and detector efficiency lower the success
    probability, ``visibility`` V leaves the heralded state (1+V)/2 |Psi+-><Psi+-| + (1-V)/2 |Psi-+><Psi-+|."""
    length_km: float = 2e-3
    p_loss_init: float = 0.0
    p_loss_length: float = 0.0          # dB / km
    speed_of_light: float = 200000.0    # km / s
    detector_efficiency: float = 1.0
    visibility: float = 1.0
    dark_count_probability: float = 0.0
    emit_duration: float = 0.0          # ns (INSTR_EMIT has no physical instruction in the reference's QPU table)
    max_num_ent_attempts: int = 100     # physical_layer.py:664

    def __post_init__(self):
        if self.dark_count_probability:
            raise NotImplementedError("heralded link: dark counts are outside the modelled (double-click, no false herald) case")

    @property
    def arm_transmission(self):
        return (1.0 - self.p_loss_init) * 10.0 ** (-self.p_loss_length * (self.length_km / 2.0) / 10.0) * self.detector_efficiency

    @property
    def success_probability(self):
        return 0.5 * self.arm_transmission ** 2

    @property
    def round_trip(self):
        """ns from emission to the herald's arrival: photon to the midpoint, outcome back."""
        return 1e9 * self.length_km / self.speed_of_light


@dataclass
class DQCSpec:
    qpus: list = field(default_factory=list)
    state4distribution: np.ndarray = None  # ket (4,) or dm (4,4); default ks.b00
    ent_delay: float = 1e9 / 182.0         # ns (examples/setup_hardware.py:18,38)
    node_separation_km: float = 2e-3       # dqc_creation.py:512-516
    handshake_wait: float = 600e3          # physical_layer.py:286
    heralded: HeraldedLinkSpec = None      # None: BlackBoxEntanglingQsourceConnection (deterministic ebits)

    def __post_init__(self):
        if self.state4distribution is None:
            self.state4distribution = B00

    @property
    def classical_delay(self):
        return 1e9 * self.node_separation_km / 200000.0  # FibreDelayModel [NS-doc]

    @property
    def ebit_is_dm(self):
        return np.asarray(self.state4distribution).size == 16

    def qpu(self, name):
        for q in self.qpus:
            if q.name == name:
                return q
        raise KeyError(name)

    @classmethod
    def uniform(cls, num_qpus, state4distribution=None, ent_delay=1e9 / 182.0, **qpu_kwargs):
        """Like ``DQC(entangling_connection_class, num_qpus, ..., qpu_class=NoisyQPU, **kw)``:
        nodes are named ``node_{i}`` (``dqc_creation.py:541-548``)."""
        heralded = qpu_kwargs.pop("heralded", None)
        return cls([QPUSpec(name=f"node_{i}", **qpu_kwargs) for i in range(num_qpus)],
                   state4distribution, ent_delay, heralded=heralded)
