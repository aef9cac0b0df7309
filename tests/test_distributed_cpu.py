"""Host-side multi-GPU logic on CPU: two processes over gloo (world_size 2).

What is checked is the part of path sharding that does not need a GPU:
contiguous chunk-aligned shards, and the ordered all-gather + tree sum that makes the
all-reduced accumulators bit-identical to a single-process reduction over all chunks."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases  # noqa: F401  (puts the package on sys.path through conftest)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import importlib
        importlib.import_module("montecarlo-risk-engine_b200")
        from mcre import runtime as RT
        assert RT.dist_info() == (rank, world)
        chunk, n_paths = 4096, 8 * 4096
        begin, count = RT.shard_range(n_paths, chunk)
        assert begin == rank * 4 * chunk and count == 4 * chunk
        # per-chunk partial sums of a fake accumulator with values spanning many magnitudes
        g = torch.Generator().manual_seed(7)
        partial = torch.randn(8, 33, dtype=torch.float64, generator=g) * torch.logspace(-8, 8, 33, dtype=torch.float64)
        mine = RT.tree_sum([partial[c] for c in range(begin // chunk, (begin + count) // chunk)])
        total = RT.all_reduce_tree(mine)
        single = RT.tree_sum([partial[c] for c in range(8)])
        ret[rank] = (bool(torch.equal(total, single)), total.numpy().tobytes())
        # integer histogram all-reduce (PFE radix select): exact and order independent
        h = torch.arange(256, dtype=torch.int64) * (rank + 1)
        dist.all_reduce(h)
        assert torch.equal(h, torch.arange(256, dtype=torch.int64) * 3)
        # ragged tail: the last rank gets the remainder, nobody gets a negative count
        b2, c2 = RT.shard_range(5 * chunk + 100, chunk)
        assert (b2, c2) == ((0, 3 * chunk) if rank == 0 else (3 * chunk, 2 * chunk + 100))
        b3, c3 = RT.shard_range(100, chunk)
        assert (b3, c3) == ((0, 100) if rank == 0 else (chunk, 0)) or c3 == 0
    finally:
        dist.destroy_process_group()


def test_sharded_tree_reduction_is_rank_count_independent():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert ret[0][0] and ret[1][0], "sharded all-reduce differs from the single-process tree sum"
    assert ret[0][1] == ret[1][1], "ranks disagree on the reduced accumulators"


def test_shard_range_partitions_the_path_range():
    import importlib
    importlib.import_module("montecarlo-risk-engine_b200")
    from mcre import runtime as RT
    for n_paths in (1, 100, 4096, 4097, 1 << 20, (1 << 20) + 12345):
        for world in (1, 2, 4, 8):
            covered = 0
            for r in range(world):
                b, c = RT.shard_range(n_paths, 4096, r, world)
                assert b % 4096 == 0 and c >= 0
                if c:
                    assert b == covered
                covered += c
            assert covered == n_paths


def _level_tree_sum(vals, group=32):
    """Python mirror of csrc/util.cu:mcre_tree_reduce (level-by-level evaluation of the binary-counter tree)."""
    def counter_fold(v, carry):
        stack = []
        for c, x in enumerate(v):
            idx = c + 1
            while idx % 2 == 0:
                x = stack.pop() + x
                idx //= 2
            stack.append(x)
        s = carry
        while stack:
            x = stack.pop()
            s = x if s is None else x + s
        return s

    cur, carry = list(vals), None
    while True:
        n = len(cur)
        last = n < 2 * group
        n_full = 0 if last else n // group
        nxt = []
        for g in range(n_full):
            v = cur[g * group:(g + 1) * group]
            w = 1
            while w < group:
                for i in range(0, group, 2 * w):
                    v[i] = v[i] + v[i + w]
                w *= 2
            nxt.append(v[0])
        if last or n % group:
            carry = counter_fold(cur[n_full * group:], carry)
        if last:
            return carry
        cur = nxt


def test_level_tree_reduction_equals_the_binary_counter_tree():
    """The device reduction evaluates the summation tree 32 elements per level; the tree itself (and so every
    bit of the result) must stay the binary-counter tree that RT.tree_sum defines for any chunk count."""
    import importlib
    importlib.import_module("montecarlo-risk-engine_b200")
    from mcre import runtime as RT
    rng = np.random.default_rng(5)
    for n in list(range(1, 140)) + [255, 256, 257, 1000, 1023, 1024, 1025, 2047, 2048, 2049, 4096, 4097, 5000, 33000]:
        vals = list(rng.standard_normal(n) * 10.0 ** rng.integers(-8, 8, n))
        a, b = RT.tree_sum(vals), _level_tree_sum(vals)
        assert a == b, (n, a, b)
