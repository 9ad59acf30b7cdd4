"""Small sharded run for ncu (launch under torchrun with 2+ ranks): QFT on LOCAL + log2(world) qubits, a few circuits in a
row so that pull passes (loads through a qubit remap from peer shards over NVLink) appear in the launch list."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from qvm_b200 import circuits  # noqa: E402
from qvm_b200.dist import ShardedState  # noqa: E402

local = int(sys.argv[1]) if len(sys.argv) > 1 else 28
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = local + world.bit_length() - 1
st = ShardedState(n, dist, device=lr)
st.set_zero_state()
gates = circuits.qft_circuit(range(n))
for _ in range(runs):
    st.apply_gates(gates, fuse=True)
nrm = st.norm2()
if rank == 0:
    print("norm2", nrm, "steps", st.steps, "peer steps", st.peer_steps, "peer GB", st.peer_bytes / 1e9, "peer s", st.peer_seconds)
st.close()
dist.destroy_process_group()
