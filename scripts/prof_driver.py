"""Small profiling driver (run under ncu on the GPU box): one fused QFT-n run and a few unfused gate
passes on low / middle / high qubits."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qvm_b200 import circuits, gates as G, qvm  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
mode = sys.argv[2] if len(sys.argv) > 2 else "all"
vec = qvm.DeviceVector(1 << n)
vec.set_zero_state()
if mode in ("all", "unfused"):
    for q in (0, 3, 12, 20, n - 1):
        vec.apply_matrix(G.gate_matrix("H"), (q,))
    vec.apply_matrix(G.gate_matrix("CNOT"), (n - 1, 2))
    vec.apply_matrix(G.gate_matrix("CPHASE", [0.3]), (5, n - 2))
if mode in ("all", "fused"):
    tape = qvm.Tape(n, circuits.qft_circuit(range(n)), fuse=True)
    print(tape.describe())
    vec.run_tape(tape)
vec.synchronize()
print("norm2", vec.norm2())
