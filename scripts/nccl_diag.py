#!/usr/bin/env python
"""What does the node's fabric give the output all-gather?  (run under torchrun, N ranks)
  * nvidia-smi topo -m (rank 0), NCCL's own transport lines (NCCL_DEBUG=INFO in the environment)
  * ncclAllGather of the bench's payload (24 x 512 x 512 x 3 fp32 = 75.5 MB per rank) with NOTHING else running: ms, bus GB/s
  * cudaMemcpyPeer bandwidth GPU0 -> GPU1 (rank 0, plain torch copy between two visible devices)"""
import os, subprocess, sys, time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
if rank == 0:
    print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout, flush=True)
src = torch.rand(24, 512, 512, 3, device=dev)
dst = torch.empty((world * 24, 512, 512, 3), device=dev)
for dt, name in ((torch.float32, "fp32 75.5 MB/rank"), (torch.float16, "fp16 37.7 MB/rank")):
    s, d = src.to(dt), dst.to(dt)
    for _ in range(3):
        dist.all_gather_into_tensor(d, s)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10):
        dist.all_gather_into_tensor(d, s)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    nbytes = s.numel() * s.element_size()
    if rank == 0:
        print(f"[diag] all_gather {name} x{world}: {ms:.3f} ms -> algbw {world * nbytes / ms / 1e6:.0f} GB/s, busbw {(world - 1) * nbytes / ms / 1e6:.0f} GB/s", flush=True)
t = torch.ones(64 << 20, device=dev)
for _ in range(3):
    dist.all_reduce(t)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(10):
    dist.all_reduce(t)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
if rank == 0:
    print(f"[diag] all_reduce 256 MB x{world}: {ms:.3f} ms -> busbw {2 * (world - 1) / world * t.numel() * 4 / ms / 1e6:.0f} GB/s", flush=True)
dist.barrier()
if rank == 0 and torch.cuda.device_count() > 1:
    a = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
    b = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:1")
    print("[diag] can_device_access_peer(0,1):", torch.cuda.can_device_access_peer(0, 1), flush=True)
    for _ in range(2):
        b.copy_(a)
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    t0 = time.perf_counter()
    for _ in range(10):
        b.copy_(a)
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    dt = (time.perf_counter() - t0) / 10
    print(f"[diag] cudaMemcpyPeer 0->1 256 MB: {dt * 1e3:.3f} ms -> {a.numel() / dt / 1e9:.0f} GB/s", flush=True)
dist.barrier()
dist.destroy_process_group()
