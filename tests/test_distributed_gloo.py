"""World-size-2 gloo test (CPU) of the multi-GPU host logic: batch sharding with
no data-path collective, optional all-gather, push-to-shared-volume all-reduce.
The compute stand-in on CPU is the oracle (the CUDA kernels need a GPU)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'torch-interpol_b200'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import oracle
        from interpol_b200 import distributed as D

        def pull(vol, grid):
            return torch.from_numpy(oracle.grid_pull(vol.numpy(), grid.numpy(), [3], [3], 1))

        def push(img, grid, shape):
            return torch.from_numpy(oracle.grid_push(img.numpy(), grid.numpy(), shape, [3], [3], 1))

        g = torch.Generator().manual_seed(0)
        B = 5                                            # uneven split: 3 + 2
        vol = torch.randn([B, 2, 6, 7, 8], generator=g, dtype=torch.float64)
        grid = torch.rand([B, 4, 5, 6, 3], generator=g, dtype=torch.float64) * 8 - 1
        lo, hi = D.shard_bounds(B, world, rank)
        assert (lo, hi) == ((0, 3) if rank == 0 else (3, 5))
        assert D.shard_batch(vol).shape[0] == hi - lo
        local = D.sharded(pull, vol, grid)
        assert local.shape[0] == hi - lo
        full = D.sharded(pull, vol, grid, gather=True)
        ref = pull(vol, grid)
        assert torch.allclose(full, ref, atol=1e-12)
        assert torch.allclose(local, ref[lo:hi], atol=1e-12)
        # a broadcast (batch 1) operand is not sliced
        one = D.sharded(pull, vol[:1], grid, gather=True)
        assert torch.allclose(one, pull(vol[:1].expand(B, -1, -1, -1, -1), grid), atol=1e-12)
        # push to a shared volume: each rank splats its own points, all-reduce(SUM)
        img = torch.randn([1, 2, 4, 5, 6], generator=g, dtype=torch.float64)
        pts = torch.rand([world, 4, 5, 6, 3], generator=g, dtype=torch.float64) * 8 - 1
        shared = D.push_to_shared(push, img, pts[rank:rank + 1], [6, 7, 8])
        want = sum(push(img, pts[r:r + 1], [6, 7, 8]) for r in range(world))
        assert torch.allclose(shared, want, atol=1e-12)
        q.put((rank, 'ok'))
    except Exception as e:          # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_sharding_and_collectives_gloo_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, 'ok'), (1, 'ok')], res


def test_shard_bounds_cover_the_batch():
    from interpol_b200.distributed import shard_bounds
    for batch in (0, 1, 5, 8, 64):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(batch, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and b >= a
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
