"""Generates tests/golden/bench_circuits.json from the reference's benchmark inputs (bench/*.quil).
Run in the build container only (reads /root/reference); the JSON is committed so that nothing on the GPU
box needs the reference tree.  qaoa_8q.quil (two 256x256 DEFGATEs, ~5 MB of text) is stored as its two
matrices in bench_qaoa_8q.npz."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from qvm_b200.quil import GateApp, Measure, parse_quil_file  # noqa: E402

REF = "/root/reference/bench"
out = {"_source": "quil-lang/qvm v1.18.0 bench/*.quil (inputs only; the reference publishes no timings)"}
for name in ("20H", "25H", "5x4x25", "entangle-25"):
    prog = parse_quil_file(os.path.join(REF, name + ".quil"))
    ins = []
    for i in prog.instructions:
        if isinstance(i, GateApp):
            ins.append(["G", i.name, list(i.params), list(i.qubits)])
        elif isinstance(i, Measure):
            ins.append(["M", i.qubit, list(i.target) if i.target else None])
    out[name] = {"source": f"bench/{name}.quil", "n_qubits": prog.qubits_needed(), "instructions": ins}
json.dump(out, open(os.path.join(HERE, "bench_circuits.json"), "w"), separators=(",", ":"))
prog = parse_quil_file(os.path.join(REF, "qaoa_8q.quil"))
np.savez_compressed(os.path.join(HERE, "bench_qaoa_8q.npz"),
                    UB=prog.gate_defs["UB-0"].matrix(), UC=prog.gate_defs["UC-0"].matrix())
print({k: len(v["instructions"]) for k, v in out.items() if k != "_source"})
