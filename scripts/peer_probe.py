"""Development aid (>= 2 GPUs, torchrun): does torch symmetric memory give peer-mapped buffers here, and how fast are pushes?"""
import os, sys, time
sys.path.insert(0, ".")
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
g = dist.group.WORLD
try:
    symm_mem.enable_symm_mem_for_group(g.group_name)
except Exception as e:
    print("enable:", repr(e))
nbytes = 32 << 20
buf = symm_mem.empty(2 * world * nbytes, dtype=torch.uint8, device=dev)
hdl = symm_mem.rendezvous(buf, g.group_name)
print(rank, "buf ptr", hex(buf.data_ptr()), "hdl", type(hdl).__name__, [a for a in dir(hdl) if not a.startswith("_")][:30], flush=True)
peers = [hdl.get_buffer(p, (2, world, nbytes), torch.uint8) for p in range(world)]
print(rank, "peer ptrs", [hex(t.data_ptr()) for t in peers], [str(t.device) for t in peers], flush=True)
src = torch.full((nbytes,), rank + 1, dtype=torch.uint8, device=dev)
for p in range(world):
    peers[p][0, rank].copy_(src, non_blocking=True)
hdl.barrier(channel=0)
torch.cuda.synchronize()
mine = buf.view(2, world, nbytes)
print(rank, "received", [int(mine[0, r, 12345]) for r in range(world)], flush=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    for p in range(world):
        peers[p][1, rank].copy_(src, non_blocking=True)
    hdl.barrier(channel=1)
e1.record(); torch.cuda.synchronize()
print(rank, f"push {nbytes >> 20} MB to {world} ranks + barrier: {e0.elapsed_time(e1) / 10:.3f} ms", flush=True)
dist.destroy_process_group()
