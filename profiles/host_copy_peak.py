"""What the host side of a box sustains in pinned host -> device copies with N ranks at once, for ordinary pinned memory
and for write-combined pinned memory (cudaHostAllocWriteCombined: not snooped through the CPU caches, meant for buffers
the CPU only writes sequentially and the GPU reads).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/host_copy_peak.py

Rank 0 prints one line per variant: GB/s per GPU (slowest rank) of 218 MB copies (the e2e step's H2D byte count)."""
import ctypes
import os
import time

import numpy as np
import torch
import torch.distributed as dist


def wc_pinned(nbytes):
    rt = ctypes.CDLL("libcudart.so.12")
    ptr = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(ptr), ctypes.c_size_t(nbytes), ctypes.c_uint(0x04))  # cudaHostAllocWriteCombined
    assert rc == 0, rc
    arr = np.ctypeslib.as_array((ctypes.c_uint8 * nbytes).from_address(ptr.value))
    return torch.from_numpy(arr), ptr


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nbytes = 217_793_536
    dst = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    variants = {"pinned (torch pin_memory)": torch.empty(nbytes, dtype=torch.uint8).pin_memory()}
    try:
        variants["pinned, write-combined (cudaHostAllocWriteCombined)"], _keep = wc_pinned(nbytes)
    except Exception as exc:  # noqa: BLE001
        if rank == 0:
            print("write-combined allocation failed:", exc)
    for name, src in variants.items():
        src[::4096] = 1  # touch every page from this rank
        assert src.is_pinned(), name
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        best = None
        for _ in range(3):
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for _ in range(10):
                dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 10
            best = dt if best is None else min(best, dt)
        t = torch.tensor([best], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"N={world}  {name:55s} {nbytes / float(t.item()) / 1e9:6.1f} GB/s per GPU (slowest rank)", flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
