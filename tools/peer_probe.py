"""Minimal 2-rank probe of the peer-memory path: k_phi_rhs storing into the other rank's rhs buffer.
torchrun --nproc-per-node 2 tools/peer_probe.py"""
import os, sys
os.environ["DOTS_PEER"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{rank}"))
from dots_socp_b200 import synth
from dots_socp_b200.engine import Engine
geo, _ = synth.example("icosphere2")
eng = Engine(7, geo, leaf_size=8)
print(rank, "peers", eng.peers, eng.peer_error, flush=True)
if eng.peers:
    eng.scale_z(2.0)
    eng.iterate(3, write_z=True)
    torch.cuda.synchronize()
    print(rank, "iterated; phi max", float(eng.slab["phi"].data.abs().max()), flush=True)
dist.barrier(); dist.destroy_process_group()
