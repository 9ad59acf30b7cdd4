"""Print the schedules one rank of a sharded state runs for consecutive circuits (no GPU, no state: schedule inspection
through the test emulator's entry points).  usage: show_sharded_schedule.py N_TOTAL WORLD [qft|random] [absorb] [pull] [runs]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
from qvm_b200 import circuits  # noqa: E402

n = int(sys.argv[1])
world = int(sys.argv[2])
kind = sys.argv[3] if len(sys.argv) > 3 else "qft"
absorb = int(sys.argv[4]) if len(sys.argv) > 4 else 0
pull = int(sys.argv[5]) if len(sys.argv) > 5 else 1
runs = int(sys.argv[6]) if len(sys.argv) > 6 else 3
circ = circuits.qft_circuit(range(n)) if kind == "qft" else circuits.random_layers(n, 10, seed=0)
emu = helpers.emulator()
emu.qvtest_shard_compile.restype = C.c_void_p
emu.qvtest_set_remap_pull(pull)
ks, qf, mf = helpers.flatten_circuit(circ)
p = lambda a: a.ctypes.data_as(C.c_void_p)
l2p = np.arange(n, dtype=np.int32)
for r in range(runs):
    err = C.create_string_buffer(512)
    t = C.c_void_p(emu.qvtest_shard_compile(n, world, 0, len(circ), p(ks), p(qf), p(mf), 1, 12, absorb, p(l2p), err, 512))
    if not t:
        raise SystemExit(err.value.decode())
    buf = C.create_string_buffer(1 << 16)
    emu.qvtest_shard_describe(t, buf, len(buf))
    print(f"--- circuit {r}\n{buf.value.decode()}")
    emu.qvtest_shard_l2p(t, p(l2p))
    emu.qvtest_shard_free(t)
