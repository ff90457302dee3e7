"""Stand-in for the NetSquid *names* the reference's front end imports (``compat/netsquid_shim``).

NetSquid 1.1.7 is a closed cp39 wheel.  The reference's pure-Python front end
(qasm2ast -> ast2dqc_circuit -> partitioner -> compilers) only needs the names below to import and run
unmodified; ``install()`` registers them in ``sys.modules`` -- when no real NetSquid is importable -- so that
``compat.frontend.run_qasm`` can run that front end in-process and hand its ``qpu_op_dict`` to the executor and
the CUDA engine.  No state arithmetic happens here: every amplitude is computed by libdqcengine.

Operator semantics follow SURVEY.md Appendix A ([NS-doc] rows): ``Operator.ctrl`` has the
control on the first index, ``conj`` is the Hermitian conjugate, ``^`` is kron and
``create_rotation_op(t, n) = cos(t/2) I - i sin(t/2) n.sigma``.
"""
import sys
import types

import numpy as np


class Operator:
    def __init__(self, name, matrix):
        self.name = name
        self.arr = np.array(matrix, dtype=complex)

    @property
    def num_qubits(self):
        return int(np.log2(self.arr.shape[0]))

    @property
    def ctrl(self):
        d = self.arr.shape[0]
        out = np.eye(2 * d, dtype=complex)
        out[d:, d:] = self.arr
        return Operator("C" + self.name, out)

    @property
    def conj(self):
        return Operator(self.name + "_dagger", self.arr.conj().T)

    def __xor__(self, other):
        return Operator(self.name + "^" + other.name, np.kron(self.arr, other.arr))

    def __repr__(self):
        return f"Operator({self.name})"


_s2 = 1 / np.sqrt(2)
I = Operator("I", [[1, 0], [0, 1]])
X = Operator("X", [[0, 1], [1, 0]])
Y = Operator("Y", [[0, -1j], [1j, 0]])
Z = Operator("Z", [[1, 0], [0, -1]])
H = Operator("H", [[_s2, _s2], [_s2, -_s2]])
S = Operator("S", [[1, 0], [0, 1j]])
T = Operator("T", [[1, 0], [0, np.exp(1j * np.pi / 4)]])
CNOT = Operator("CNOT", X.ctrl.arr)
CZ = Operator("CZ", Z.ctrl.arr)
CS = Operator("CS", S.ctrl.arr)
SWAP = Operator("SWAP", [[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]])


def create_rotation_op(angle, rotation_axis=(0, 0, 1), conjugate=False):
    n = np.array(rotation_axis, dtype=float)
    n = n / np.linalg.norm(n)
    m = np.cos(angle / 2) * I.arr - 1j * np.sin(angle / 2) * (
        n[0] * X.arr + n[1] * Y.arr + n[2] * Z.arr
    )
    if conjugate:
        m = m.conj().T
    return Operator("R", m)


class Instruction:
    def __init__(self, name, operator=None, num_positions=None):
        self.name = name
        self.operator = operator
        if num_positions is None and operator is not None:
            num_positions = operator.num_qubits
        self.num_positions = num_positions

    def __repr__(self):
        return f"Instruction({self.name})"


class IGate(Instruction):
    def __init__(self, name, operator=None, num_positions=-1):
        super().__init__(name, operator, None if operator is not None else num_positions)


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Anything:
    """Base class / callable placeholder for DES plumbing we never execute."""

    def __init__(self, *a, **k):
        pass

    def __init_subclass__(cls, **k):
        pass


class QuantumProcessor(_Anything):
    """Only stores the memory layout the partitioner reads (quantum_processors.py:83-123)."""

    def __init__(self, name, num_positions=1, mem_noise_models=None, phys_instructions=None,
                 mem_pos_types=None, fallback_to_nonphysical=False, properties=None, **kw):
        self.name = name
        self.num_positions = num_positions
        self.phys_instructions = phys_instructions
        self.mem_noise_models = mem_noise_models

    def add_composite_instruction(self, *a, **k):
        pass


class PhysicalInstruction(_Anything):
    def __init__(self, instruction, duration=0, **kw):
        self.instruction = instruction
        self.duration = duration
        self.kw = kw


class QuantumErrorModel(_Anything):
    def __init__(self, **kw):
        self.properties = {}

    def add_property(self, name, value, **kw):
        self.properties[name] = value


class DepolarNoiseModel(QuantumErrorModel):
    def __init__(self, depolar_rate=0, time_independent=False, **kw):
        super().__init__()
        self.depolar_rate = depolar_rate
        self.time_independent = time_independent


def install():
    if "netsquid" in sys.modules and getattr(sys.modules["netsquid"], "__stub__", False):
        return
    try:  # real pyparsing is absent in the image; pip vendors 3.1.0
        import pyparsing  # noqa: F401
    except ModuleNotFoundError:
        import pip._vendor.pyparsing as _pp
        sys.modules["pyparsing"] = _pp

    ins = dict(
        Instruction=Instruction, IGate=IGate,
        INSTR_INIT=Instruction("init"), INSTR_MEASURE=Instruction("measure", num_positions=1),
        INSTR_DISCARD=Instruction("discard"), INSTR_EMIT=Instruction("emit"),
        INSTR_H=IGate("h_gate", H), INSTR_X=IGate("x_gate", X), INSTR_Y=IGate("y_gate", Y),
        INSTR_Z=IGate("z_gate", Z), INSTR_S=IGate("s_gate", S), INSTR_T=IGate("t_gate", T),
        INSTR_I=IGate("i_gate", I),
        INSTR_CNOT=IGate("cnot_gate", CNOT), INSTR_CZ=IGate("cz_gate", CZ),
        INSTR_CS=IGate("cs_gate", CS), INSTR_SWAP=IGate("swap_gate", SWAP),
        INSTR_CX=IGate("cnot_gate", CNOT),
    )
    instructions = _mod("netsquid.components.instructions", **ins)
    qprocessor = _mod("netsquid.components.qprocessor", QuantumProcessor=QuantumProcessor,
                      PhysicalInstruction=PhysicalInstruction)
    qprogram = _mod("netsquid.components.qprogram", QuantumProgram=_Anything)
    qerr = _mod("netsquid.components.models.qerrormodels", QuantumErrorModel=QuantumErrorModel,
                DepolarNoiseModel=DepolarNoiseModel)
    delay = _mod("netsquid.components.models.delaymodels", FibreDelayModel=_Anything)
    models = _mod("netsquid.components.models", qerrormodels=qerr, delaymodels=delay,
                  DepolarNoiseModel=DepolarNoiseModel, FibreDelayModel=_Anything)
    cchannel = _mod("netsquid.components.cchannel", ClassicalChannel=_Anything)
    components = _mod(
        "netsquid.components", instructions=instructions, qprocessor=qprocessor,
        qprogram=qprogram, models=models, cchannel=cchannel, QuantumProgram=_Anything,
        ClassicalChannel=_Anything, QuantumChannel=_Anything, QSource=_Anything,
        SourceStatus=types.SimpleNamespace(EXTERNAL=0, INTERNAL=1, OFF=2),
        FibreDelayModel=_Anything, QuantumProcessor=QuantumProcessor,
        PhysicalInstruction=PhysicalInstruction)

    operators = _mod("netsquid.qubits.operators", Operator=Operator, I=I, X=X, Y=Y, Z=Z, H=H,
                     S=S, T=T, CNOT=CNOT, CZ=CZ, CS=CS, SWAP=SWAP,
                     create_rotation_op=create_rotation_op)
    s0 = np.array([[1], [0]], dtype=complex)
    s1 = np.array([[0], [1]], dtype=complex)
    k = np.kron
    ketstates = _mod(
        "netsquid.qubits.ketstates", s0=s0, s1=s1, h0=(s0 + s1) * _s2, h1=(s0 - s1) * _s2,
        b00=(k(s0, s0) + k(s1, s1)) * _s2, b01=(k(s0, s1) + k(s1, s0)) * _s2,
        b10=(k(s0, s0) - k(s1, s1)) * _s2, b11=(k(s0, s1) - k(s1, s0)) * _s2,
        BellIndex=types.SimpleNamespace(PHI_PLUS=0, PSI_PLUS=1, PSI_MINUS=2, PHI_MINUS=3))
    ketutil = _mod("netsquid.qubits.ketutil", ket2dm=lambda v: np.asarray(v) @ np.asarray(v).conj().T)
    qubitapi = _mod("netsquid.qubits.qubitapi")
    dmtools = _mod("netsquid.qubits.dmtools", DenseDMRepr=_Anything)
    dmutil = _mod("netsquid.qubits.dmutil", partialtrace=None, reorder_dm=None)
    qformalism = _mod("netsquid.qubits.qformalism",
                      QFormalism=types.SimpleNamespace(KET="ket", DM="dm", SPARSEDM="sparsedm", STAB="stab", GSLC="gslc"),
                      set_qstate_formalism=lambda f: None)
    qubit = _mod("netsquid.qubits.qubit", Qubit=_Anything)
    qubits = _mod("netsquid.qubits", operators=operators, ketstates=ketstates, ketutil=ketutil,
                  qubitapi=qubitapi, dmtools=dmtools, dmutil=dmutil, qformalism=qformalism,
                  qubit=qubit, StateSampler=_Anything, QFormalism=qformalism.QFormalism)

    nodes_conn = _mod("netsquid.nodes.connections", Connection=_Anything, DirectConnection=_Anything)
    nodes = _mod("netsquid.nodes", Node=_Anything, Network=_Anything, DirectConnection=_Anything,
                 Connection=_Anything, connections=nodes_conn)
    protocol = _mod("netsquid.protocols.protocol", Signals=types.SimpleNamespace(), Protocol=_Anything)
    nodeprotocols = _mod("netsquid.protocols.nodeprotocols", NodeProtocol=_Anything,
                         LocalProtocol=_Anything)
    protocols = _mod("netsquid.protocols", NodeProtocol=_Anything, LocalProtocol=_Anything,
                     Signals=types.SimpleNamespace(), protocol=protocol, nodeprotocols=nodeprotocols)
    cmap = _mod("netsquid.util.constrainedmap", ValueConstraint=_Anything)
    dcol = _mod("netsquid.util.datacollector", DataCollector=_Anything)
    util = _mod("netsquid.util", constrainedmap=cmap, datacollector=dcol)
    ns = _mod("netsquid", components=components, qubits=qubits, nodes=nodes, protocols=protocols,
              util=util, set_qstate_formalism=lambda f: None, sim_reset=lambda: None)
    ns.__stub__ = True
    _mod("pydynaa", Entity=_Anything, EventType=_Anything, EventExpression=_Anything)
    hc = _mod("netsquid_physlayer.heralded_connection", MiddleHeraldedConnection=_Anything)
    _mod("netsquid_physlayer", heralded_connection=hc)
    ref_src = "/root/reference/src"
    if ref_src not in sys.path:
        sys.path.insert(0, ref_src)
