"""Multi-GPU data parallelism for the forward rendering path.

Every render (voxel, pose) is independent -- the forward graph has no cross-sample op -- so the path shards over
the batch axis with NO data-path collective: rank g renders items [lo, hi) of the batch with its own replica of
the weights.  The only collective is the all-gather of the output image batch the north_star asks for (plus a
one-off broadcast of the weights / the turntable voxel at start-up).  One process per GPU, torch.distributed
(NCCL on GPUs; the same code runs on gloo/CPU tensors for the host-logic tests).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice of `n_items` owned by `rank`; the first n_items % world ranks get one extra item."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_gather_images(local: torch.Tensor, n_total: Optional[int] = None, group=None) -> torch.Tensor:
    """[b_local,H,W,C] on every rank -> [n_total,H,W,C] on every rank (rank-major order == batch order for
    shard_bounds).  Ragged shards are padded to the largest shard for the collective and trimmed afterwards."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    if n_total is None:
        cnt = torch.tensor([local.shape[0]], device=local.device, dtype=torch.int64)
        dist.all_reduce(cnt, group=group)
        n_total = int(cnt.item())
    sizes = [shard_bounds(n_total, world, r) for r in range(world)]
    bmax = max(hi - lo for lo, hi in sizes)
    if all(hi - lo == bmax for lo, hi in sizes):
        out = torch.empty((world * bmax,) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    pad = torch.zeros((bmax,) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
    pad[: local.shape[0]] = local
    buf = torch.empty((world * bmax,) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
    dist.all_gather_into_tensor(buf, pad, group=group)
    parts = [buf[r * bmax: r * bmax + (hi - lo)] for r, (lo, hi) in enumerate(sizes)]
    return torch.cat(parts, dim=0)


class PeerImageGather:
    """All-gather of the per-rank image shards WITHOUT SM time: every rank owns a full [world*b, ...] buffer, exported to
    its peers over CUDA IPC at start-up; each step a rank writes its shard into every peer's buffer with plain
    device-to-device copies (copy engines over NVLink / NVSwitch) on a side stream.

    An alternative to ncclAllGather that leaves all 148 SMs to the persistent convolution kernels of the next step (NCCL's
    kernels hold a few SMs while they run).  Measured at 2 GPUs it makes no difference (41.6 vs 41.3 ms/step: the NCCL
    gather already hides completely behind the step), so `bench.py` keeps NCCL as the default and offers this as
    `--gather peer`.  Single node only (IPC); construction raises if peer access or IPC is unavailable, and the caller
    falls back to `all_gather_images` (NCCL)."""

    def __init__(self, shard_shape, dtype=torch.float32, device=None, group=None):
        if not dist.is_initialized():
            raise RuntimeError("PeerImageGather needs an initialised process group")
        from torch.multiprocessing.reductions import reduce_tensor
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.b = int(shard_shape[0])
        self.full = torch.empty((self.world * self.b,) + tuple(shard_shape[1:]), device=self.device, dtype=dtype)
        handles = [None] * self.world
        dist.all_gather_object(handles, reduce_tensor(self.full), group=group)
        self.peers = []
        for r, (rebuild, rargs) in enumerate(handles):
            if r == self.rank:
                self.peers.append(self.full)
                continue
            t = rebuild(*rargs)                                  # aliases rank r's buffer (cudaIpcOpenMemHandle)
            if not torch.cuda.can_device_access_peer(self.device.index, t.device.index):
                raise RuntimeError(f"no peer access {self.device} -> {t.device}")
            self.peers.append(t)
        self.stream = torch.cuda.Stream(device=self.device)
        self.done = torch.cuda.Event()
        self.done.record(torch.cuda.current_stream(self.device))

    def gather_async(self, local: torch.Tensor, ready: Optional[torch.cuda.Event] = None) -> torch.cuda.Event:
        """Queue the copies of `local` ([b, ...], this rank's shard) into every rank's buffer on the side stream, after
        `ready` (default: everything queued so far on the current stream).  Returns the event that marks this rank's
        outgoing copies complete; incoming shards are complete once every rank's event has fired (`fence()`)."""
        if ready is None:
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(self.device))
        lo = self.rank * self.b
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            for k in range(self.world):                          # own copy first, then peers rank+1, rank+2, ...: no hot spot
                self.peers[(self.rank + k) % self.world][lo:lo + self.b].copy_(local, non_blocking=True)
            self.done = torch.cuda.Event()
            self.done.record(self.stream)
        return self.done

    def fence(self):
        """Block until every rank's outgoing copies have landed (so `self.full` is complete everywhere)."""
        self.done.synchronize()
        dist.barrier(group=self.group)

    def gather(self, local: torch.Tensor) -> torch.Tensor:
        self.gather_async(local)
        self.fence()
        return self.full

    def close(self):
        """Drop the peer mappings before the owners free their buffers."""
        self.fence()
        self.peers = [self.full]
        dist.barrier(group=self.group)


def broadcast_weight_dict(weights: Optional[Dict[str, np.ndarray]], src: int = 0, device="cpu", group=None):
    """Replicate a weight dict from `src` to every rank (names first, then one flat fp32 buffer)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return weights
    meta = [[(k, tuple(np.asarray(v).shape)) for k, v in sorted(weights.items())]] if dist.get_rank(group) == src else [None]
    dist.broadcast_object_list(meta, src=src, group=group)
    names = meta[0]
    total = sum(int(np.prod(s)) for _, s in names)
    flat = torch.empty(total, dtype=torch.float32, device=device)
    if dist.get_rank(group) == src:
        flat.copy_(torch.cat([torch.as_tensor(np.asarray(weights[k], np.float32)).reshape(-1) for k, _ in names]))
    dist.broadcast(flat, src=src, group=group)
    out, off = {}, 0
    flat_cpu = flat.cpu().numpy()
    for k, s in names:
        n = int(np.prod(s))
        out[k] = flat_cpu[off:off + n].reshape(s)
        off += n
    return out


def turntable_poses(n_frames: int = 360, elevation: float = 60.0, radius: float = 3.3, step_deg: Optional[float] = None):
    """Poses of the demo's azimuth sweep (RenderNet_demo.py:133: 0..355 step 5; BASELINE config 5: 0..359 step 1)."""
    step = step_deg if step_deg is not None else 360.0 / n_frames
    az = np.arange(n_frames, dtype=np.float64) * step
    return np.stack([az * math.pi / 180.0, np.full(n_frames, (90 - elevation) * math.pi / 180),
                     np.full(n_frames, 3.3 / radius)], axis=1).astype(np.float32)
