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


def _hybrid_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from scail_b200.parallel import HybridParallel
        hp = HybridParallel()
        assert hp.cp_size == world // 2 and hp.branch == rank // hp.cp_size
        if hp.cp_size > 1:
            assert hp.cp.size == hp.cp_size and hp.cp.rank == rank % hp.cp_size
            # context-parallel token sharding inside this branch's half only
            full = torch.arange(2 * 8 * 3, dtype=torch.float32).view(1, 16, 3) + 1000 * hp.branch
            assert torch.equal(hp.cp.gather_tokens(hp.cp.shard_tokens(full), 16), full)
        else:
            assert hp.cp is None
        # partner exchange: (uncond, cond) order on every rank, partner = same CP rank in the other half
        v = torch.full((1, 4), float(10 * hp.branch + rank % hp.cp_size))
        both = hp.gather_branches(v)
        assert both.shape == (2, 4)
        assert torch.equal(both[0], torch.full((4,), float(rank % hp.cp_size)))
        assert torch.equal(both[1], torch.full((4,), float(10 + rank % hp.cp_size)))
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def _run_hybrid(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_hybrid_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


def test_hybrid_cfg_parallel_world2():
    _run_hybrid(2)


def test_hybrid_cfg_x_cp_world4():
    _run_hybrid(4)
