"""baseline/cpu_port.c is the CPU BASELINE bench.py times (AVX2/FMA + OpenMP port of the reference's kernels); it is
not the checker, but a baseline that computed something else would time the wrong thing: hold it to the oracle."""
import numpy as np
import pytest

import helpers as H
from qvm_b200 import circuits


@pytest.mark.parametrize("threads", [1, 4])
def test_cpu_port_matches_oracle(threads):
    from baseline import cpu_port as P
    rng = np.random.default_rng(1)
    n = 12
    circ = H.random_circuit(n, 80, rng, max_dense=4) + circuits.qft_circuit(range(n))
    psi = H.rand_state(n)
    ref = H.run_oracle(psi.copy(), circ)
    for m, q in circ:
        P.apply_gate(psi, m, q, threads)
    H.assert_close(psi, ref)


def test_cpu_port_uses_all_cores_despite_omp_env(monkeypatch):
    import os
    from baseline import cpu_port as P
    monkeypatch.setenv("OMP_NUM_THREADS", "1")      # what torchrun exports
    assert P.all_cores() == len(os.sched_getaffinity(0))
    psi = P.zero_state(10, P.all_cores())
    assert psi[0] == 1.0 and np.count_nonzero(psi) == 1
