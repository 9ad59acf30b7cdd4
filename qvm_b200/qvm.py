"""Host-side mirror of the reference QVM API for the hot path, over libqvmcuda.

SBCL is not available in the build image, so the Lisp shim in lisp/ can only be
written, not exercised; this module drives the SAME C ABI in the order the Lisp
generics would (SURVEY.md section 3 call stacks) under the reference's names:

    make_qvm / make_density_qvm      qvm:make-qvm (src/qvm.lisp:150-164), make-density-qvm (src/density-qvm.lisp:60-71)
    load_program / run / run_program src/classical-memory-mixin.lisp:114-159, src/execution.lisp:13-59
    apply_gate_to_state              src/apply-gate.lisp:106-212
    measure / measure_all            src/measurement.lisp:87-162
    amplitudes                       qvm::amplitudes (src/qvm.lisp:63-69)

Random draws stay on the host (numpy's MT19937 RandomState, the reference uses
mt19937 as well); the library only ever receives uniforms.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib as L
from . import gates as G
from .quil import Declare, GateApp, Halt, Measure, Program, Reset, parse_quil

# ---- src/config.lisp:41-58 switches that keep their meaning ---------------------------------
compile_before_running = False      # *compile-before-running*
fuse_gates_during_compilation = True  # *fuse-gates-during-compilation*
compile_measure_chains = True       # *compile-measure-chains*


class DeviceVector:
    """A device-resident CFLONUM vector: what `allocate-vector` on a CUDA-ALLOCATION returns
    (src/allocator.lisp:47-62).  Zero-initialised; freed by close() / the finalizer."""

    def __init__(self, length: int, device: int = 0):
        self.length = int(length)
        self.device = device
        h = C.c_void_p()
        L.check(L.lib().qvmcuda_state_create(self.length, device, C.byref(h)))
        self.handle = h

    def close(self):
        if getattr(self, "handle", None):
            L.lib().qvmcuda_state_destroy(self.handle)
            self.handle = None

    __del__ = close

    # -- state protocol ------------------------------------------------------------------
    def download(self, offset: int = 0, count: Optional[int] = None) -> np.ndarray:
        count = self.length - offset if count is None else count
        out = np.empty(count, dtype=np.complex128)
        L.check(L.lib().qvmcuda_download(self.handle, L.ptr(out), offset, count))
        return out

    def download_into(self, dst: np.ndarray, offset: int = 0) -> None:
        """Straight into caller-owned memory (a POSIX shared-memory mapping, a memory-mapped dump): no intermediate buffer."""
        if dst.dtype != np.complex128 or not dst.flags["C_CONTIGUOUS"] or not dst.flags["WRITEABLE"]:
            raise ValueError("destination must be a writable contiguous complex128 array")
        L.check(L.lib().qvmcuda_download(self.handle, L.ptr(dst), offset, dst.size))

    def upload(self, data, offset: int = 0) -> None:
        a = np.ascontiguousarray(data, dtype=np.complex128)
        L.check(L.lib().qvmcuda_upload(self.handle, L.ptr(a), offset, a.size))

    def set_zero_state(self): L.check(L.lib().qvmcuda_set_zero_state(self.handle))
    def set_basis_state(self, b: int): L.check(L.lib().qvmcuda_set_basis_state(self.handle, int(b)))
    def synchronize(self): L.check(L.lib().qvmcuda_synchronize(self.handle))
    def set_stream(self, cuda_stream: int): L.check(L.lib().qvmcuda_state_set_stream(self.handle, int(cuda_stream)))

    def copy_from(self, other: "DeviceVector"):
        L.check(L.lib().qvmcuda_copy(self.handle, other.handle))

    # -- operator API -----------------------------------------------------------------------
    def apply_matrix(self, matrix, qubits: Sequence[int]) -> None:
        """qvm:apply-matrix-operator; QUBITS in Quil argument order."""
        ks, qf, mf = L.flatten_gates([(matrix, tuple(qubits))])
        L.check(L.lib().qvmcuda_apply_matrix(self.handle, int(ks[0]), L.ptr(qf), L.ptr(mf)))

    def apply_gates(self, gates, fuse: bool = True, absorb_swaps: bool = False) -> None:
        if not gates:
            return
        ks, qf, mf = L.flatten_gates(gates)
        flags = (L.FUSE if fuse else 0) | (L.ABSORB_SWAPS if absorb_swaps else 0)
        L.check(L.lib().qvmcuda_apply_gates(self.handle, len(gates), L.ptr(ks), L.ptr(qf), L.ptr(mf), flags))

    def run_tape(self, tape: "Tape") -> None:
        L.check(L.lib().qvmcuda_tape_run(self.handle, tape.handle))

    # -- measurement protocol ---------------------------------------------------------------
    def prob_excited(self, q: int) -> float:
        p = C.c_double()
        L.check(L.lib().qvmcuda_prob_excited(self.handle, q, C.byref(p)))
        return p.value

    def prob_ground(self, q: int) -> float:
        p = C.c_double()
        L.check(L.lib().qvmcuda_prob_ground(self.handle, q, C.byref(p)))
        return p.value

    def norm2(self) -> float:
        p = C.c_double()
        L.check(L.lib().qvmcuda_norm2(self.handle, C.byref(p)))
        return p.value

    def probabilities(self, offset: int = 0, count: Optional[int] = None) -> np.ndarray:
        """|psi_i|^2 of COUNT basis states from OFFSET, computed on the device (PERFORM-PROBABILITIES,
        app/src/api/probabilities.lisp): 8 bytes per basis state cross PCIe instead of 16."""
        count = self.length - offset if count is None else count
        out = np.zeros(count, dtype=np.float64)
        L.check(L.lib().qvmcuda_probabilities(self.handle, L.ptr(out), int(offset), int(count)))
        return out

    def inner_product(self, other: "DeviceVector") -> complex:
        """<self|other> = sum conj(self_i) other_i (app/src/api/expectation.lisp:79-84)."""
        out = np.zeros(2, dtype=np.float64)
        L.check(L.lib().qvmcuda_inner_product(self.handle, other.handle, L.ptr(out)))
        return complex(out[0], out[1])

    def scale(self, f: float): L.check(L.lib().qvmcuda_scale(self.handle, float(f)))
    def normalize(self): L.check(L.lib().qvmcuda_normalize(self.handle))

    def collapse(self, q: int, keep_bit: int, inv_norm: float):
        L.check(L.lib().qvmcuda_collapse(self.handle, q, keep_bit, float(inv_norm)))

    def sample(self, uniforms, strict: bool = False) -> np.ndarray:
        u = np.ascontiguousarray(uniforms, dtype=np.float64)
        out = np.empty(u.size, dtype=np.uint64)
        L.check(L.lib().qvmcuda_sample(self.handle, L.ptr(u), u.size, L.ptr(out), 1 if strict else 0))
        return out

    def density_apply_ops(self, n: int, ops, fuse: bool = True):
        """A run of density operators in ONE library call (qvmcuda_density_apply_ops).
        ops = [([K_0, ...] Kraus matrices, qubits in Quil argument order)]."""
        if not ops:
            return
        ks = np.ascontiguousarray([len(q) for _, q in ops], dtype=np.int32)
        ms = np.ascontiguousarray([len(k) for k, _ in ops], dtype=np.int32)
        qf = np.ascontiguousarray([int(x) for _, q in ops for x in reversed(q)], dtype=np.int32)
        kf = np.ascontiguousarray(np.concatenate([np.asarray(m, dtype=np.complex128).ravel() for k, _ in ops for m in k])).view(np.float64)
        for k, q in ops:
            for m in k:
                if np.asarray(m).shape != (1 << len(q), 1 << len(q)):
                    raise ValueError("Kraus matrix does not match its qubit count")
        L.check(L.lib().qvmcuda_density_apply_ops(self.handle, n, len(ops), L.ptr(ks), L.ptr(qf), L.ptr(ms), L.ptr(kf),
                                                  L.FUSE if fuse else 0))

    def density_expectation(self, n: int, op_matrix) -> complex:
        """tr(Q rho), MIXED-STATE-EXPECTATION (app/src/api/expectation.lisp:91-107)."""
        q = np.ascontiguousarray(op_matrix, dtype=np.complex128)
        if q.shape != (1 << n, 1 << n):
            raise ValueError("operator matrix must be 2^n x 2^n")
        out = np.zeros(2, dtype=np.float64)
        L.check(L.lib().qvmcuda_density_expectation(self.handle, n, L.ptr(q), L.ptr(out)))
        return complex(out[0], out[1])

    def set_identity_matrix(self, n: int):
        L.check(L.lib().qvmcuda_set_identity_matrix(self.handle, n))

    def sample_total(self) -> float:
        """This vector's probability mass in the sampler's own summation order (sharded sampling)."""
        t = C.c_double(0.0)
        L.check(L.lib().qvmcuda_sample_total(self.handle, C.byref(t)))
        return float(t.value)

    def sample_shard(self, uniforms, strict: bool, base: float) -> np.ndarray:
        """Shard-local indices for draws this shard owns; base = mass of the shards in front of it."""
        u = np.ascontiguousarray(uniforms, dtype=np.float64)
        out = np.empty(u.size, dtype=np.uint64)
        L.check(L.lib().qvmcuda_sample_shard(self.handle, L.ptr(u), u.size, L.ptr(out), 1 if strict else 0, float(base)))
        return out

    # -- density ------------------------------------------------------------------------------
    def density_apply_kraus(self, n: int, kraus, qubits: Sequence[int], fuse: bool = True):
        ks = np.ascontiguousarray(np.stack([np.asarray(k, dtype=np.complex128) for k in kraus]))
        q = np.ascontiguousarray(list(reversed([int(x) for x in qubits])), dtype=np.int32)
        L.check(L.lib().qvmcuda_density_apply_kraus(self.handle, n, len(q), L.ptr(q), len(kraus), L.ptr(ks),
                                                    L.FUSE if fuse else 0))

    def density_prob_excited(self, n: int, q: int) -> float:
        p = C.c_double()
        L.check(L.lib().qvmcuda_density_prob_excited(self.handle, n, q, C.byref(p)))
        return p.value

    def density_collapse(self, n, q, keep_bit, inv_norm):
        L.check(L.lib().qvmcuda_density_collapse(self.handle, n, q, keep_bit, float(inv_norm)))

    def density_measure_discard(self, n, q):
        L.check(L.lib().qvmcuda_density_measure_discard(self.handle, n, q))

    def density_diag_probs(self, n) -> np.ndarray:
        out = np.empty(1 << n, dtype=np.float64)
        L.check(L.lib().qvmcuda_density_diag_probs(self.handle, n, L.ptr(out)))
        return out


class Tape:
    """A compiled gate sequence: what COMPILE-LOADED-PROGRAM builds once (src/qvm.lisp:166-175)."""

    def __init__(self, n_qubits: int, gates, fuse: bool = True, absorb_swaps: bool = False):
        ks, qf, mf = L.flatten_gates(gates)
        h = C.c_void_p()
        flags = (L.FUSE if fuse else 0) | (L.ABSORB_SWAPS if absorb_swaps else 0)
        L.check(L.lib().qvmcuda_tape_compile(n_qubits, len(gates), L.ptr(ks), L.ptr(qf), L.ptr(mf), flags, C.byref(h)))
        self.handle = h
        self.n_gates = len(gates)

    def info(self) -> Dict[str, int]:
        a = np.zeros(8, dtype=np.int64)
        L.check(L.lib().qvmcuda_tape_info(self.handle, L.ptr(a)))
        return {"passes": int(a[0]), "gates": int(a[1]), "atoms": int(a[2]), "tile_passes": int(a[3]),
                "generic_passes": int(a[4]), "table_bytes": int(a[5])}

    def describe(self) -> str:
        buf = C.create_string_buffer(1 << 16)
        L.check(L.lib().qvmcuda_tape_describe(self.handle, buf, len(buf)))
        return buf.value.decode()

    def close(self):
        if getattr(self, "handle", None):
            L.lib().qvmcuda_tape_destroy(self.handle)
            self.handle = None

    __del__ = close


# ---------------------------------------------------------------------------- states
class PureState:
    """PURE-STATE (src/state-representation.lisp:58-167) over a device vector."""

    def __init__(self, num_qubits: int, device: int = 0):
        self.num_qubits = num_qubits
        self.vec = DeviceVector(1 << num_qubits, device)
        self.vec.set_zero_state()           # (setf (aref amplitudes 0) 1), :92-95
        self.trial = None                   # TRIAL-AMPLITUDES, allocated on first noisy gate

    def state_elements(self) -> np.ndarray: return self.vec.download()
    def set_state_elements(self, amps): self.vec.upload(amps)
    def set_to_zero_state(self): self.vec.set_zero_state()


class DensityMatrixState:
    """DENSITY-MATRIX-STATE (src/state-representation.lisp:173-286): vec(rho), row-major, 4^n entries."""

    def __init__(self, num_qubits: int, device: int = 0):
        self.num_qubits = num_qubits
        self.vec = DeviceVector(1 << (2 * num_qubits), device)
        self.vec.set_zero_state()

    def state_elements(self) -> np.ndarray: return self.vec.download()
    def set_state_elements(self, v): self.vec.upload(v)
    def set_to_zero_state(self): self.vec.set_zero_state()

    def matrix_view(self) -> np.ndarray:
        d = 1 << self.num_qubits
        return self.state_elements().reshape(d, d)

    def measurement_probabilities(self) -> np.ndarray:
        """DENSITY-MATRIX-STATE-MEASUREMENT-PROBABILITIES :268-286."""
        return self.vec.density_diag_probs(self.num_qubits)

    def apply_ops(self, ops, fuse: bool = True) -> None:
        """A run of unitaries / Kraus channels in ONE library call (fused passes over vec(rho)), through the same
        entry point the Lisp shim's gate tape flushes into (lisp/operators.lisp FLUSH-GATE-TAPE)."""
        self.vec.density_apply_ops(self.num_qubits, [(list(g) if isinstance(g, (list, tuple)) else [g], q) for g, q in ops], fuse=fuse)


def density_gate_list(n: int, ops):
    """Translate density-matrix operations into gates on the 2n index bits of vec(rho).

    ops = [(U or [K_0, K_1, ...], qubits in Quil order)].  A unitary becomes conj(U) on the column bits
    (qubits) and U on the row bits (qubits + n), exactly the two applications of src/apply-gate.lisp:56-65.
    A Kraus list becomes ONE superoperator sum_j K_j (x) conj(K_j) on (row bits, column bits) instead of the
    reference's copy / apply / add / restore loop (src/apply-gate.lisp:79-99)."""
    out = []
    for gate, qubits in ops:
        qubits = tuple(int(q) for q in qubits)
        ghosts = tuple(q + n for q in qubits)
        if isinstance(gate, (list, tuple)) and len(gate) > 1:
            sop = sum(np.kron(np.asarray(k, dtype=np.complex128), np.conj(np.asarray(k, dtype=np.complex128))) for k in gate)
            out.append((np.ascontiguousarray(sop), ghosts + qubits))
        else:
            u = np.asarray(gate[0] if isinstance(gate, (list, tuple)) else gate, dtype=np.complex128)
            out.append((np.ascontiguousarray(np.conj(u)), qubits))
            out.append((np.ascontiguousarray(u), ghosts))
    return out


class CompiledGateApplication:
    """Stand-in for COMPILED-MATRIX / COMPILED-INLINED-MATRIX / COMPILED-PERMUTATION-GATE-APPLICATION
    (src/compile-gate.lisp:363-409): an instruction that carries its matrix and its own arguments.  The reference's
    TRANSITION hands such an instruction to APPLY-GATE-TO-STATE with QUBITS = NIL (src/transition.lisp:182-186)."""

    def __init__(self, matrix, qubits: Sequence[int]):
        self.matrix = np.ascontiguousarray(matrix, dtype=np.complex128)
        self.qubits = tuple(int(q) for q in qubits)


def apply_gate_to_state(gate, state, qubits: Optional[Sequence[int]]) -> None:
    """APPLY-GATE-TO-STATE (src/apply-gate.lisp:106-212).  GATE is a matrix, a list of Kraus matrices (a KRAUS-LIST
    superoperator) or a compiled gate application; QUBITS in Quil argument order, or None for a compiled gate
    application (the qubits then come from the instruction, as in lisp/operators.lisp)."""
    if isinstance(gate, CompiledGateApplication):
        qubits = tuple(qubits) if qubits else gate.qubits
        gate = gate.matrix
    if qubits is None:
        raise ValueError("only a compiled gate application carries its own qubits")
    if isinstance(state, PureState):
        if isinstance(gate, (list, tuple)):
            raise ValueError("a Kraus list on a pure state needs a uniform draw: use evolve_pure_state_stochastically")
        state.vec.apply_matrix(gate, qubits)
    else:
        kraus = list(gate) if isinstance(gate, (list, tuple)) else [gate]
        state.vec.density_apply_kraus(state.num_qubits, kraus, qubits)


def evolve_pure_state_stochastically(kraus_map, state: "PureState", qubits: Sequence[int], r: float) -> int:
    """%EVOLVE-PURE-STATE-STOCHASTICALLY (src/apply-gate.lisp:16-39): pick Kraus operator j with probability
    <psi|K_j^dagger K_j|psi> by inverse-transform sampling with the host-drawn uniform r, lazily: every
    candidate is applied to a scratch copy (the TRIAL-AMPLITUDES of src/state-representation.lisp:63-70) and
    its squared norm accumulated until the sum reaches r; the trial then becomes the state and is normalised.
    Returns the index of the operator applied."""
    if state.trial is None:                      # CHECK-ALLOCATE-COMPUTATION-SPACE :104-115
        state.trial = DeviceVector(state.vec.length, state.vec.device)
    summed, j = 0.0, 0
    for j, k in enumerate(kraus_map):
        state.trial.copy_from(state.vec)
        state.trial.apply_matrix(k, qubits)
        summed += state.trial.norm2()
        if summed >= r:
            break
    state.vec, state.trial = state.trial, state.vec          # (rotatef amplitudes trial-amplitudes)
    state.vec.normalize()
    return j


# ---------------------------------------------------------------------------- machines
class BaseQVM:
    def __init__(self, seed: Optional[int] = None):
        self.program: Optional[Program] = None
        self.pc = 0
        self.rng = np.random.RandomState(seed)      # MT19937, like the reference's mt19937:random
        self.registers: Dict[str, np.ndarray] = {}
        self._compiled = None

    # classical memory: only what MEASURE needs (the rest is out of the hot path)
    def _store(self, target, bit):
        if target is None:
            return
        name, off = target
        # the reference signals an error for a memory reference outside the declared region
        # (DEREFERENCE-MREF, src/classical-memory-mixin.lisp); so do we, instead of growing the register
        if name not in self.registers:
            raise KeyError(f"MEASURE into undeclared memory region {name!r}")
        if off < 0 or off >= self.registers[name].size:
            raise IndexError(f"memory reference {name}[{off}] is outside the declared region of {self.registers[name].size}")
        self.registers[name][off] = bit

    def load_program(self, program):
        if isinstance(program, str):
            program = parse_quil(program)
        self.program = program
        self.pc = 0
        self._compiled = None
        self.registers = {}                 # a new program starts from its own DECLAREs only
        for ins in program.instructions:
            if isinstance(ins, Declare):
                self.registers[ins.name] = np.zeros(ins.length, dtype=np.int64)
        return self

    def number_of_qubits(self) -> int:
        return self.state.num_qubits

    def random(self) -> float:
        return float(self.rng.random_sample())


class PureStateQVM(BaseQVM):
    """PURE-STATE-QVM (src/qvm.lisp:114-175)."""

    def __init__(self, num_qubits: int, device: int = 0, seed: Optional[int] = None):
        super().__init__(seed)
        self.state = PureState(num_qubits, device)
        self.superoperator_definitions: Dict[Tuple[str, Tuple[int, ...]], list] = {}

    def set_superoperator(self, name: str, qubits: Sequence[int], kraus) -> None:
        """SET-SUPEROPERATOR: replace gate NAME on QUBITS by a Kraus map, applied stochastically
        (src/transition.lisp:155-180, tests/state-representation-tests.lisp:55-66)."""
        G.check_kraus_ops(list(kraus))
        self.superoperator_definitions[(name, tuple(qubits))] = list(kraus)

    # qvm::amplitudes / (setf qvm::amplitudes) : the device<->host sync points
    @property
    def amplitudes(self) -> np.ndarray:
        return self.state.state_elements()

    @amplitudes.setter
    def amplitudes(self, v):
        self.state.set_state_elements(v)

    def reset_quantum_state(self):
        self.state.set_to_zero_state()

    # -- measurement.lisp:87-105 (interpreted) and compile-gate.lisp:221-254 (compiled) ----
    def measure(self, q: int) -> int:
        assert 0 <= q < self.number_of_qubits()
        p1 = self.state.vec.prob_excited(q)
        cbit = 0 if p1 == 0.0 else (1 if self.random() <= p1 else 0)
        inv = 1.0 / math.sqrt(p1) if cbit == 1 else 1.0 / math.sqrt(1.0 - p1)
        self.state.vec.collapse(q, cbit, inv)
        return cbit

    def measure_compiled(self, q: int) -> int:
        p0 = self.state.vec.prob_ground(q)      # WAVEFUNCTION-GROUND-STATE-PROBABILITY
        keep_zero = self.random() < p0
        bit = 0 if keep_zero else 1
        inv = 1.0 / math.sqrt(p0) if keep_zero else 1.0 / math.sqrt(1.0 - p0)
        self.state.vec.collapse(q, bit, inv)
        return bit

    def measure_all(self) -> List[int]:
        """MEASURE-ALL-STATE (pure) src/measurement.lisp:128-143: one uniform, C(b) > p rule, psi <- |b>."""
        b = int(self.state.vec.sample([self.random()], strict=True)[0])
        self.state.vec.set_basis_state(b)
        return [(b >> i) & 1 for i in range(self.number_of_qubits())]

    def sample_multiple(self, n_shots: int) -> np.ndarray:
        """SAMPLE-WAVEFUNCTION-MULTIPLE-TIMES src/measurement.lisp:246-288 with host-drawn uniforms."""
        return self.state.vec.sample(self.rng.random_sample(n_shots), strict=False)

    # -- run loop (src/execution.lisp:13-59, src/transition.lisp) ---------------------------------
    def _gate_list(self, instrs) -> List[Tuple[np.ndarray, Tuple[int, ...]]]:
        return [(self.program.gate_matrix(i), i.qubits) for i in instrs]

    def run(self):
        prog = self.program
        n = self.number_of_qubits()
        ins = prog.instructions
        i = 0
        compiled = compile_before_running
        while i < len(ins):
            x = ins[i]
            if isinstance(x, GateApp):
                if compiled:
                    # COMPILE-LOADED-PROGRAM: a maximal run of gates becomes one fused tape
                    # (same predicate as the interpreted path below: a gate with modifiers is never the noisy gate)
                    def noisy(g):
                        return (g.name, tuple(g.qubits)) in self.superoperator_definitions and not g.modifiers
                    j = i
                    while j < len(ins) and isinstance(ins[j], GateApp) and not noisy(ins[j]):
                        j += 1
                    if j == i:        # a noisy gate: stochastic evolution, like the interpreted path
                        evolve_pure_state_stochastically(self.superoperator_definitions[(x.name, tuple(x.qubits))],
                                                         self.state, x.qubits, self.random())
                        i += 1
                        continue
                    self.state.vec.apply_gates(self._gate_list(ins[i:j]), fuse=fuse_gates_during_compilation)
                    i = j
                    continue
                key = (x.name, tuple(x.qubits))
                if key in self.superoperator_definitions and not x.modifiers:
                    evolve_pure_state_stochastically(self.superoperator_definitions[key], self.state, x.qubits, self.random())
                else:
                    self.state.vec.apply_matrix(prog.gate_matrix(x), x.qubits)
            elif isinstance(x, Measure):
                if compiled and compile_measure_chains:
                    # COMPILE-MEASURE-CHAINS src/compile-gate.lisp:551-599
                    j = i
                    while j < len(ins) and isinstance(ins[j], Measure):
                        j += 1
                    chain = ins[i:j]
                    if len(chain) >= n and {m.qubit for m in chain} == set(range(n)):
                        bits = self.measure_all()
                        for m in chain:
                            self._store(m.target, bits[m.qubit])
                        i = j
                        continue
                bit = self.measure_compiled(x.qubit) if compiled else self.measure(x.qubit)
                self._store(x.target, bit)
            elif isinstance(x, Reset):
                if x.qubit is None:
                    self.state.set_to_zero_state()
                else:
                    # RESET q = MEASURE q, then X if 1 (src/transition.lisp:76-95)
                    if self.measure(x.qubit) == 1:
                        self.state.vec.apply_matrix(G.gate_matrix("X"), (x.qubit,))
            elif isinstance(x, Halt):
                break
            i += 1
        self.pc = i
        return self


class DensityQVM(BaseQVM):
    """DENSITY-QVM (src/density-qvm.lisp:33-190): always interpreted (:183-190)."""

    def __init__(self, num_qubits: int, device: int = 0, seed: Optional[int] = None):
        super().__init__(seed)
        self.state = DensityMatrixState(num_qubits, device)
        self.noisy_gate_definitions: Dict[Tuple[str, Tuple[int, ...]], list] = {}
        self.readout_povms: Dict[int, Tuple[float, float, float, float]] = {}
        self.fuse_gates = True
        self._tape: list = []

    @property
    def amplitudes(self) -> np.ndarray:
        return self.state.state_elements()

    @amplitudes.setter
    def amplitudes(self, v):
        self.state.set_state_elements(v)

    def reset_quantum_state(self):
        self.state.set_to_zero_state()

    def set_noisy_gate(self, name: str, qubits: Sequence[int], kraus) -> None:
        """SET-NOISY-GATE src/density-qvm.lisp:73-82."""
        G.check_kraus_ops(list(kraus))
        self.noisy_gate_definitions[(name, tuple(qubits))] = list(kraus)

    def set_readout_povm(self, qubit: int, povm) -> None:
        self.readout_povms[qubit] = tuple(povm)

    def measure(self, q: int) -> int:
        n = self.number_of_qubits()
        p1 = self.state.vec.density_prob_excited(n, q)
        cbit = 0 if p1 == 0.0 else (1 if self.random() <= p1 else 0)
        inv = 1.0 / p1 if cbit == 1 else 1.0 / (1.0 - p1)
        self.state.vec.density_collapse(n, q, cbit, inv)
        return cbit

    def measure_all(self) -> List[int]:
        """NAIVE-MEASURE-ALL src/measurement.lisp:153-162, then POVM bit flips (density-qvm.lisp:169-180)."""
        n = self.number_of_qubits()
        bits = [0] * n
        for q in range(n - 1, -1, -1):
            bits[q] = self.measure(q)
        return [self._perturb(q, b) for q, b in enumerate(bits)]

    def _perturb(self, q: int, bit: int) -> int:
        povm = self.readout_povms.get(q)
        if povm is None:
            return bit
        p00, p01, p10, p11 = povm          # PERTURB-MEASUREMENT src/channel-qvm.lisp:185-195; p(observed|actual)
        r = self.random()
        if bit == 0:
            return 0 if r <= p00 else 1
        return 0 if r <= p01 else 1

    def flush_gate_tape(self):
        """What the Lisp shim does when something needs rho (lisp/operators.lisp FLUSH-GATE-TAPE): all pending gate /
        channel transitions (src/density-qvm.lisp:125-137) go to the device in ONE qvmcuda_density_apply_ops call."""
        if self._tape:
            self.state.vec.density_apply_ops(self.number_of_qubits(), self._tape, fuse=self.fuse_gates)
            self._tape = []

    def run(self):
        prog = self.program
        n = self.number_of_qubits()
        self._tape = []
        for x in prog.instructions:
            if isinstance(x, GateApp):
                key = (x.name, tuple(x.qubits))
                if key in self.noisy_gate_definitions and not x.modifiers:
                    self._tape.append((self.noisy_gate_definitions[key], tuple(x.qubits)))
                else:
                    self._tape.append(([prog.gate_matrix(x)], tuple(x.qubits)))
                continue
            self.flush_gate_tape()
            if isinstance(x, Measure):
                if x.target is None:
                    self.state.vec.density_measure_discard(n, x.qubit)
                else:
                    self._store(x.target, self._perturb(x.qubit, self.measure(x.qubit)))
            elif isinstance(x, Reset):
                if x.qubit is None:
                    self.state.set_to_zero_state()
                else:
                    raise NotImplementedError("RESET q on a density matrix is outside the hot path")
            elif isinstance(x, Halt):
                break
        self.flush_gate_tape()
        return self


class UnitaryQVM(BaseQVM):
    """UNITARY-QVM (src/unitary-qvm.lisp:30-137): computes the unitary matrix of a program.  The state is a 4^n vector
    holding the matrix in COLUMN-major order, initialised to the identity; gates act on the low n index bits (the row
    index), i.e. it is the pure-state path on 2n bits.  No measurement."""

    def __init__(self, num_qubits: int, device: int = 0):
        super().__init__(None)
        self.num_qubits = num_qubits
        self.state = PureState(2 * num_qubits, device)
        self.state.vec.set_identity_matrix(num_qubits)

    def number_of_qubits(self) -> int:
        return self.num_qubits

    def reset_quantum_state(self):
        self.state.vec.set_identity_matrix(self.num_qubits)

    def measure(self, q):
        raise RuntimeError("MEASURE unsupported in unitary calculation.")

    def measure_all(self):
        raise RuntimeError("MEASURE unsupported in unitary calculation.")

    def run(self):
        gates = []
        for x in self.program.instructions:
            if isinstance(x, GateApp):
                gates.append((self.program.gate_matrix(x), x.qubits))
            elif isinstance(x, (Measure, Reset)):
                raise RuntimeError("MEASURE / RESET unsupported in unitary calculation.")
        self.state.vec.apply_gates(gates, fuse=fuse_gates_during_compilation)
        return self

    def underlying_matrix(self) -> np.ndarray:
        """UNITARY-QVM-UNDERLYING-MATRIX: column-major storage -> (row, column) array."""
        d = 1 << self.num_qubits
        return self.state.vec.download().reshape(d, d).T.copy()


def parsed_program_unitary_matrix(program, num_qubits: int, device: int = 0) -> np.ndarray:
    """PARSED-PROGRAM-UNITARY-MATRIX (src/unitary-qvm.lisp:128-137)."""
    q = UnitaryQVM(num_qubits, device)
    q.load_program(program)
    q.run()
    m = q.underlying_matrix()
    q.state.vec.close()
    return m


def mixed_state_expectation(qvm: "DensityQVM", op_matrix) -> complex:
    """MIXED-STATE-EXPECTATION (app/src/api/expectation.lisp:91-107): tr(Q rho) reduced on the device."""
    qvm.flush_gate_tape()
    return qvm.state.vec.density_expectation(qvm.number_of_qubits(), op_matrix)


def wavefunction_octets(amplitudes) -> bytes:
    """The :wavefunction reply body (app/src/handle-request.lisp:135-153, WRITE-COMPLEX-DOUBLE-FLOAT-AS-BINARY
    app/src/utilities.lisp:64-78): every amplitude as two big-endian IEEE-754 doubles, 16 octets each."""
    a = np.ascontiguousarray(amplitudes, dtype=np.complex128)
    return a.view(np.float64).astype(">f8").tobytes()


def probabilities_octets(probabilities) -> bytes:
    """The :probabilities reply body (app/src/handle-request.lisp:155-176, WRITE-DOUBLE-FLOAT-AS-BINARY
    app/src/utilities.lisp:80-87): big-endian doubles, 8 octets each."""
    return np.ascontiguousarray(probabilities, dtype=np.float64).astype(">f8").tobytes()


UNUSED_QUBIT = None      # the reference's :unused-qubit marker (app/src/api/multishot-measure.lisp:66-69)


def relabel_multishot_qubits(qubits: Sequence[int], num_qubits: int, relabeling: Optional[Sequence[int]]):
    """The relabeling step of %PERFORM-MULTISHOT-MEASURE (app/src/api/multishot-measure.lisp:62-77): qubit q becomes
    its position in RELABELING (or :unused-qubit when absent), and the register grows to hold the largest one."""
    qubits = list(qubits)
    if any((not isinstance(q, (int, np.integer))) or q < 0 for q in qubits):
        raise ValueError("qubits must be non-negative integers")
    if relabeling is not None:
        rel = [int(x) for x in relabeling]
        qubits = [rel.index(q) if q in rel else UNUSED_QUBIT for q in qubits]
        used = [q for q in qubits if q is not UNUSED_QUBIT]
        num_qubits = max(num_qubits, 1 + max(used, default=-1))
    return qubits, num_qubits


def multishot_bits(basis_states, qubits: Sequence[Optional[int]]) -> List[List[int]]:
    """What PARALLEL-MEASURE collects per trial (app/src/api/multishot-measure.lisp:13-38): bit q of the measured basis
    state for every requested qubit, 0 for :unused-qubit, in the order of QUBITS."""
    b = np.asarray(basis_states, dtype=np.uint64)
    cols = [np.zeros(b.size, dtype=np.int64) if q is UNUSED_QUBIT else ((b >> np.uint64(q)) & np.uint64(1)).astype(np.int64)
            for q in qubits]
    return np.stack(cols, axis=1).tolist() if cols else [[] for _ in range(b.size)]


def perform_multishot_measure(quil, num_qubits: int, qubits: Sequence[int], num_trials: int,
                              relabeling: Optional[Sequence[int]] = None, device: int = 0, seed: Optional[int] = None):
    """%PERFORM-MULTISHOT-MEASURE on a pure state (app/src/api/multishot-measure.lisp:50-110).  The reference prepares
    the state once and then, PER TRIAL, restores a 2^n copy and collapses it with MEASURE-ALL (or qubit by qubit);
    here the prepared state never leaves the device and all NUM-TRIALS outcomes are drawn from it in ONE sampler call
    (one read of the state + NUM-TRIALS descents) with MEASURE-ALL's decision rule (strict, src/measurement.lisp:128-143).
    Measuring a subset of the qubits is the marginal of the same draw (the alternative the reference's own XXX comment
    debates, :27-31)."""
    if not qubits or num_trials == 0:
        return []
    qubits, num_qubits = relabel_multishot_qubits(qubits, num_qubits, relabeling)
    if any(q is not UNUSED_QUBIT and q >= num_qubits for q in qubits):
        raise ValueError(f"The provided qubits {qubits} to a multishot measure are out of range for the given QVM, "
                         f"which only has {num_qubits} qubits.")
    qvm = make_qvm(num_qubits, device=device, seed=seed)
    qvm.load_program(quil)
    qvm.run()
    u = qvm.rng.random_sample(num_trials)
    outcomes = qvm.state.vec.sample(u, strict=True)
    qvm.state.vec.close()
    return multishot_bits(outcomes, qubits)


def pure_state_expectation(qvm: "PureStateQVM", prepared: DeviceVector, op, first_time: bool = False) -> complex:
    """PURE-STATE-EXPECTATION (app/src/api/expectation.lisp:78-91): restore the prepared state, run the operator
    program OP on it and return <prepared | OP prepared>.  Everything stays on the device."""
    if not first_time:
        qvm.state.vec.copy_from(prepared)
    qvm.load_program(op)
    qvm.run()
    return prepared.inner_product(qvm.state.vec)


def perform_expectation(state_prep, operators, num_qubits: int, device: int = 0, seed: Optional[int] = None) -> List[float]:
    """PERFORM-EXPECTATION on a pure state (app/src/api/expectation.lisp:38-76): run STATE-PREP once, keep a device copy
    of the prepared wavefunction, then one expectation value per operator program.  The imaginary part must vanish
    to 1e-14 like in the reference (:73)."""
    qvm = make_qvm(num_qubits, device=device, seed=seed)
    qvm.load_program(state_prep)
    qvm.run()
    prepared = DeviceVector(1 << num_qubits, device)
    prepared.copy_from(qvm.state.vec)
    out = []
    first = True
    for op in operators:
        e = pure_state_expectation(qvm, prepared, op, first_time=first)
        first = False
        if abs(e.imag) >= 1e-14:
            raise ValueError(f"expectation value has an imaginary part: {e}")
        out.append(e.real)
    prepared.close()
    qvm.state.vec.close()
    return out


def make_qvm(num_qubits: int, device: int = 0, seed: Optional[int] = None) -> PureStateQVM:
    return PureStateQVM(num_qubits, device, seed)


def make_density_qvm(num_qubits: int, device: int = 0, seed: Optional[int] = None) -> DensityQVM:
    return DensityQVM(num_qubits, device, seed)


def run_program(num_qubits: int, program, device: int = 0, seed: Optional[int] = None) -> PureStateQVM:
    """RUN-PROGRAM src/execution.lisp:46-59."""
    qvm = make_qvm(num_qubits, device, seed)
    qvm.load_program(program)
    return qvm.run()
