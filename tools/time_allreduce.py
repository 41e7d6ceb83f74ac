"""2+ GPU (torchrun): 60.8 MB fp32 / 30.4 MB bf16 all-reduce: NCCL vs torch symmetric-memory multimem / two-shot kernels."""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
gname = dist.group.WORLD.group_name
n = 15_200_000
for dtype in (torch.float32, torch.bfloat16):
    plain = torch.ones(n, dtype=dtype, device=dev)
    sym = symm_mem.empty(n, dtype=dtype, device=dev)
    sym.fill_(1)
    symm_mem.rendezvous(sym, gname)
    variants = {"nccl": lambda: dist.all_reduce(plain)}
    # NCCL user-buffer registration: the tensor comes from ncclMemAlloc (torch.cuda.MemPool over the backend's allocator)
    # and the pool is registered with the communicator -> zero-copy / NVLS all-reduce that needs few SMs
    try:
        backend = dist.group.WORLD._get_backend(dev)
        pool = torch.cuda.MemPool(backend.mem_allocator)
        with torch.cuda.use_mem_pool(pool):
            reg = torch.ones(n, dtype=dtype, device=dev)
        backend.register_mem_pool(pool)
        variants["nccl registered"] = lambda: dist.all_reduce(reg)
    except Exception as ex:
        if rank == 0:
            print("registered pool unavailable:", repr(ex)[:200], flush=True)
    for opname in ("multimem_all_reduce_", "two_shot_all_reduce_", "one_shot_all_reduce"):
        op = getattr(torch.ops.symm_mem, opname, None)
        if op is not None:
            variants[opname] = (lambda op=op: op(sym, "sum", gname))
    for name, fn in variants.items():
        try:
            for _ in range(3):
                fn()
            torch.cuda.synchronize(); dist.barrier()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(20):
                fn()
            e.record(); torch.cuda.synchronize()
            if rank == 0:
                us = s.elapsed_time(e) * 1e3 / 20
                print(f"{str(dtype):16s} {name:24s} {us:8.1f} us  algbw {n * plain.element_size() / us / 1e3:7.1f} GB/s", flush=True)
        except Exception as ex:
            if rank == 0:
                print(dtype, name, "failed:", repr(ex)[:150], flush=True)
        dist.barrier()
dist.destroy_process_group()
