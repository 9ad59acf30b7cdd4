"""dqvm's deterministic program list (dqvm/tests/program-tests.lisp:61-113): every program is run on the single-process
reference path and on the distributed one from the DEBUG WAVEFUNCTION psi_i = i (:14-19) and the two must agree.
Here: oracle vs the sharded schedules (2 and 4 emulated ranks over 4 qubits are too small for a tile, so the list runs
on 4 logical qubits embedded in an 8-qubit register sharded over 2 / 4 / 8 ranks, both exchange modes)."""
import numpy as np
import pytest

from helpers import assert_close, run_emulator_sharded, run_oracle, unpermute
from qvm_b200.quil import parse_quil

PROGRAMS = ["I 0", "X 1", "Y 2", "Z 3", "I 3; X 2; Y 1; Z 0", "SWAP 0 1", "SWAP 0 2",
            "CNOT 0 1", "CNOT 1 0", "CNOT 0 2", "CNOT 2 0", "CNOT 1 2", "CNOT 2 1", "CNOT 0 3", "CNOT 3 0", "CNOT 1 3",
            "CNOT 3 1", "CNOT 2 3", "CNOT 3 2",
            "CCNOT 0 1 2", "CCNOT 0 2 1", "CCNOT 1 0 2", "CCNOT 2 0 1", "CCNOT 1 2 0", "CCNOT 2 1 0",
            "I 0; S 1", "S 0; T 1", "I 0; T 1", "T 0; I 1", "CZ 0 1", "CZ 1 0", "CZ 0 2", "CZ 2 0", "CZ 1 2", "CZ 2 1",
            "ISWAP 0 1", "ISWAP 1 0", "ISWAP 0 2", "ISWAP 2 0", "ISWAP 1 2", "ISWAP 2 1",
            "CCNOT 0 1 2; X 1; CCNOT 0 2 1; Y 2; CCNOT 1 0 2; H 0; CCNOT 2 0 1; Z 1; CCNOT 1 2 0; H 2; CCNOT 2 1 0"]
# (the list's "H 0; H 1; H 2; RESET 1; RESET 2" needs measurement outcomes: covered by tests/test_dist_gloo.py)


def _circuit(quil, qubit_map):
    p = parse_quil(quil.replace("; ", "\n"))
    return [(p.gate_matrix(i), tuple(qubit_map[q] for q in i.qubits)) for i in p.instructions if type(i).__name__ == "GateApp"]


@pytest.mark.parametrize("world,pull", [(2, False), (4, True), (8, True), (8, False)])
def test_deterministic_programs(world, pull):
    n = 8
    # logical qubits 0..3 of the dqvm programs sit on the TOP qubits, i.e. on and next to the rank bits
    qubit_map = {0: 7, 1: 6, 2: 5, 3: 4}
    for prog in PROGRAMS:
        circ = _circuit(prog, qubit_map)
        psi = np.arange(1 << n).astype(np.complex128)      # the debug wavefunction
        psi[0] = 0
        ref = run_oracle(psi.copy(), circ)
        _, _, desc, l2p = run_emulator_sharded(psi, n, world, circ, fuse=True, tile_bits=4, remap_pull=pull)
        assert_close(unpermute(psi, l2p), ref, rel=1e-12, abs_=1e-11), prog
