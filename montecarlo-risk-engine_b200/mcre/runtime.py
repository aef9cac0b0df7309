"""Process-level runtime: one process per GPU, torch only as plumbing (device memory,
streams, torch.distributed)."""
from __future__ import annotations

import os

import torch

from mcre.binding import McreError


_device = None


def compute_device():
    """The CUDA device of this process (LOCAL_RANK under torchrun).  No CPU fallback."""
    global _device
    if _device is not None:      # (torch.cuda.is_available() costs 0.25 ms per call: thousands of calls per large book)
        return _device
    if not torch.cuda.is_available():
        raise McreError(
            "No CUDA device visible: the Monte Carlo hot path runs only on the GPU "
            "(hand-written sm_100a kernels); there is no CPU fallback.")
    idx = int(os.environ.get("LOCAL_RANK", "0")) % torch.cuda.device_count()
    torch.cuda.set_device(idx)
    _device = torch.device("cuda", idx)
    return _device


#: bytes read back device -> host through to_host() (bench.py: d2h_bytes_per_step of the end-to-end leg)
d2h_bytes = 0


def to_host(t):
    """Device tensor -> numpy array (synchronises on the tensor's stream); counts the bytes."""
    global d2h_bytes
    d2h_bytes += t.numel() * t.element_size()
    return t.cpu().numpy()


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def dist_info():
    """(rank, world_size) of the path-sharding group (1 process per GPU)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_paths, chunk, rank=None, world=None):
    """Contiguous, chunk-aligned slice of the global path range owned by `rank`.

    Chunks (fixed blocks of `chunk` global path ids) are dealt out in contiguous runs so
    that, when the chunk count per rank is a power of two, each rank's local tree sum is
    a node of the global reduction tree (GPU-count-independent bits)."""
    if rank is None:
        rank, world = dist_info()
    n_chunks = (n_paths + chunk - 1) // chunk
    per = (n_chunks + world - 1) // world
    c0, c1 = min(rank * per, n_chunks), min((rank + 1) * per, n_chunks)
    begin = c0 * chunk
    end = min(c1 * chunk, n_paths)
    return begin, max(end - begin, 0)


def tree_sum(parts):
    """Pairwise (binary-counter) sum of a list of equally shaped arrays, matching the
    device tree reduction's ordering across ranks."""
    stack = []
    for i, v in enumerate(parts):
        idx = i + 1
        while idx % 2 == 0:
            v = stack.pop() + v
            idx //= 2
        stack.append(v)
    total = None
    while stack:
        v = stack.pop()
        total = v if total is None else v + total
    return total


def all_reduce_tree(t):
    """Sum a small device tensor over ranks in a fixed tree order (all-gather + ordered
    sum; the accumulators are KB-sized, latency bound)."""
    import torch.distributed as dist
    rank, world = dist_info()
    if world == 1:
        return t
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t.contiguous())
    return tree_sum(parts)
