"""world_size-2 gloo tests (CPU) of the context-parallel host logic: token sharding, the batch-major K/V
gather-buffer layout and the in-place all-gather — everything in scail_b200.parallel that is not a kernel."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from scail_b200.parallel import ContextParallel
        cp = ContextParallel()
        B, N, d = 2, 12, 4
        full = torch.arange(B * N * d, dtype=torch.float32).view(B, N, d)
        loc = cp.shard_tokens(full)
        assert loc.shape == (B, N // world, d)
        assert torch.equal(loc, full[:, rank * (N // world):(rank + 1) * (N // world)])
        back = cp.gather_tokens(loc, N)
        assert torch.equal(back, full)
        # K/V buffer: every rank fills its slot with (batch, global token) codes; after the gather batch b must be
        # the contiguous token-ordered [N, 2d] matrix
        n = N // world
        kv = cp.kv_buffer(B, n, d, "cpu", torch.float32)
        kv.fill_(-1)
        for b in range(B):
            tok = torch.arange(rank * n, (rank + 1) * n, dtype=torch.float32)
            kv[b, rank] = (1000 * b + tok)[:, None].expand(n, 2 * d)
        for w in cp.gather_kv(kv, async_op=True):
            w.wait()
        for b in range(B):
            want = (1000 * b + torch.arange(N, dtype=torch.float32))[:, None].expand(N, 2 * d)
            assert torch.equal(kv[b].reshape(N, 2 * d), want)
        cos = torch.arange(N * 3, dtype=torch.float32).view(N, 3)
        c, s = cp.rope_slice(cos, -cos, n)
        assert torch.equal(c, cos[rank * n:(rank + 1) * n]) and torch.equal(s, -c)
        try:
            cp.local_len(13)
            ok = False
        except ValueError:
            ok = True
        assert ok
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_context_parallel_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
