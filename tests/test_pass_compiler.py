"""The pass compiler (qvm_b200/csrc/qv_jit_gen.cpp, qv_jit.cpp) without a GPU.

* the text it generates, compiled for the HOST (g++ -DQVJ_HOST) and run by the test emulator in place of its micro-op
  interpreter, reproduces the oracle on the reference's circuits (QFT of examples/qft.lisp, bench/*.quil, random mixes,
  density channels, sharded schedules with fused pull remaps);
* NVRTC turns the same text into sm_100a cubins without spills (nvcc cross-compiles here; nothing is launched);
* the cache key depends on the structure of a pass, not on its angles or on where its qubits sit.
"""
import ctypes as C

import numpy as np
import pytest

import helpers as H
from qvm_b200 import _lib, circuits, gates as G, qvm


@pytest.fixture()
def jit_host():
    before = H.emulator_jit_host(True)
    yield lambda: H.emulator_jit_host(True) - before
    H.emulator_jit_host(False)


@pytest.mark.parametrize("n", [12, 14, 17])
def test_compiled_qft_matches_oracle(jit_host, n):
    circ = circuits.qft_circuit(range(n))
    psi = H.rand_state(n)
    ref = H.run_oracle(psi.copy(), circ)
    steps, _, _ = H.run_emulator(psi, n, circ)
    assert jit_host() == steps or jit_host() >= 1      # store-permutation-only passes have no rounds to compile
    H.assert_close(psi, ref)


@pytest.mark.parametrize("seed", range(8))
def test_compiled_random_circuits_match_oracle(jit_host, seed):
    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(12, 16))
    circ = H.random_circuit(n, 70, rng)
    psi = H.rand_state(n, seed)
    ref = H.run_oracle(psi.copy(), circ)
    H.run_emulator(psi, n, circ)
    assert jit_host() >= 1
    H.assert_close(psi, ref)


def test_compiled_wide_rounds_match_oracle(jit_host):
    """16 amplitudes per thread (reg_bits = 4)."""
    rng = np.random.default_rng(7)
    circ = circuits.qft_circuit(range(14)) + H.random_circuit(14, 40, rng)
    psi = H.rand_state(14, 3)
    ref = H.run_oracle(psi.copy(), circ)
    H.run_emulator(psi, 14, circ, reg_bits=4)
    assert jit_host() >= 1
    H.assert_close(psi, ref)


@pytest.mark.parametrize("name", ["5x4x25", "20H"])
def test_compiled_bench_files(jit_host, name):
    circ, _, n = H.load_bench_circuit(name)
    if n > 20:
        pytest.skip("too large for the CPU emulator")
    psi = np.zeros(1 << n, dtype=np.complex128)
    psi[0] = 1.0
    ref = H.run_oracle(psi.copy(), circ)
    H.run_emulator(psi, n, circ)
    H.assert_close(psi, ref)


def test_compiled_density_channels(jit_host):
    """Noisy QAOA on vec(rho), 6 qubits (12 index bits): superoperator passes through compiled rounds."""
    n = 6
    circ = circuits.qaoa_maxcut_circuit(n, circuits.line_graph(n))
    dep = G.depolarizing_kraus_map(0.01)
    ops = []
    for m, q in circ:
        ops.append((m, q))
        ops.extend((dep, (qq,)) for qq in q)
    gl = qvm.density_gate_list(n, ops)
    rho = np.zeros(1 << (2 * n), dtype=np.complex128)
    rho[0] = 1.0
    ref = H.run_oracle(rho.copy(), gl)
    H.run_emulator(rho, 2 * n, gl)
    assert jit_host() >= 1
    H.assert_close(rho, ref)
    assert abs(ref.reshape(1 << n, 1 << n).trace() - 1.0) < 1e-12


@pytest.mark.parametrize("world", [2, 4])
def test_compiled_sharded_pull_passes(jit_host, world):
    """Fused pull remaps (loads through the remap, stores into the alternate buffer) with compiled rounds."""
    n = 16
    rng = np.random.default_rng(11 + world)
    circ = circuits.qft_circuit(range(n)) + H.random_circuit(n, 30, rng)
    psi = H.rand_state(n, 5)
    ref = H.run_oracle(psi.copy(), circ)
    _, peer_steps, _, l2p = H.run_emulator_sharded(psi, n, world, circ, remap_pull=True)
    assert peer_steps >= 1 and jit_host() >= 1
    H.assert_close(H.unpermute(psi, l2p), ref)


def _precompile(tape):
    ne, nk = C.c_int(), C.c_int()
    log = C.create_string_buffer(1 << 20)
    _lib.check(_lib.lib().qvmcuda_tape_jit_precompile(tape.handle, C.byref(ne), C.byref(nk), log, len(log)))
    return ne.value, nk.value, log.value.decode()


def _signatures(tape):
    sigs = []
    buf = C.create_string_buffer(1 << 20)
    for step in range(tape.info()["passes"]):
        sig = C.c_uint64()
        if _lib.lib().qvmcuda_tape_jit_source(tape.handle, step, buf, len(buf), C.byref(sig)) == 0:
            sigs.append(sig.value)
        else:
            sigs.append(None)
    return sigs


def test_nvrtc_compiles_passes_without_heavy_spills(tmp_path, monkeypatch):
    monkeypatch.setenv("QVMCUDA_JIT_CACHE", str(tmp_path))     # read once per process: only effective in a fresh one
    rng = np.random.default_rng(3)
    for circ, n in ((circuits.qft_circuit(range(16)), 16), (H.random_circuit(14, 60, rng), 14),
                    (circuits.random_layer_circuit(16, 3, 0), 16)):
        tape = qvm.Tape(n, circ, fuse=True)
        ne, nk, log = _precompile(tape)
        if "libnvrtc" in log and "not found" in log:
            pytest.skip("NVRTC not available in this container")
        assert ne >= 1 and nk == ne, log[-2000:]
        import re
        for line in log.splitlines():
            m = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m:       # a handful of spilled uniform registers is tolerated; a spilling hot loop is not
                assert int(m.group(1)) <= 256 and int(m.group(2)) <= 256, line


def test_cache_key_is_structural():
    """Same circuit shape with other angles, or moved to other (high) qubits: same kernels."""
    def layer(qs, theta):
        c = []
        for i, q in enumerate(qs):
            c.append((G.gate_matrix("RX", [theta + i]), (q,)))
            c.append((G.gate_matrix("RZ", [2 * theta + i]), (q,)))
        for a, b in zip(qs[:-1], qs[1:]):
            c.append((G.gate_matrix("CPHASE", [theta * (a + 1)]), (a, b)))
        return c
    a = _signatures(qvm.Tape(24, layer(list(range(8, 16)), 0.3), fuse=True))
    b = _signatures(qvm.Tape(24, layer(list(range(8, 16)), 1.7), fuse=True))
    c = _signatures(qvm.Tape(24, layer(list(range(14, 22)), 0.9), fuse=True))
    assert a == b == c and a[0] is not None
    # (table offsets and the presence of a store permutation are literals -- measured faster than reading them at run time --
    # so passes whose phase tables differ in size, like the QFT's three heavy passes, get their own kernels)
