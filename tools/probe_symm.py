"""Probe: does torch symmetric memory give a multicast (NVLS) pointer on this box?  Run under torchrun, 2+ GPUs."""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
t = symm_mem.empty(1 << 20, dtype=torch.float32, device=dev)
hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
print(rank, "multicast_ptr", hex(hdl.multicast_ptr) if hdl.multicast_ptr else 0, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal_pad", len(hdl.signal_pad_ptrs), flush=True)
t.fill_(rank + 1.0)
hdl.barrier()
peer = hdl.get_buffer((rank + 1) % world, (1 << 20,), torch.float32)
print(rank, "peer value", float(peer[0]), flush=True)
if hdl.multicast_ptr:
    out = torch.empty(1 << 20, dtype=torch.float32, device=dev)
    try:
        torch.ops.symm_mem.multimem_all_reduce_(t, "sum", dist.group.WORLD.group_name)
        torch.cuda.synchronize()
        print(rank, "multimem_all_reduce_ ->", float(t[0]), flush=True)
    except Exception as e:
        print(rank, "multimem_all_reduce_ failed:", repr(e)[:200], flush=True)
dist.barrier()
dist.destroy_process_group()
