"""CPU (gloo, world_size 2): host-side logic of the multi-GPU path -- video sharding and the token all-gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from videotgb_b200.distributed import TokenAllGather, shard_videos


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        videos, nframe = 6, 3
        lo, hi = shard_videos(videos, rank, world)
        g = torch.Generator().manual_seed(0)
        all_tokens = torch.randn(videos * nframe, 32, 768, generator=g)         # every rank can rebuild the truth
        local = all_tokens[lo * nframe: hi * nframe]
        ex = TokenAllGather()
        ex.exchange(local)
        got = ex.wait()
        ok = torch.equal(got, all_tokens.to(torch.bfloat16)) and ex.bytes_per_rank == local.numel() * 2
        # a second exchange re-uses the buffer
        ex.exchange(local * 2)
        ok = ok and torch.equal(ex.wait(), (all_tokens * 2).to(torch.bfloat16))
        q.put((rank, bool(ok), (lo, hi)))
    except Exception as e:                                  # noqa: BLE001  -- surface the failure instead of hanging the parent
        q.put((rank, False, repr(e)))
    finally:
        dist.destroy_process_group()


def test_token_allgather_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] and res[1][1]
    assert res[0][2] == (0, 3) and res[1][2] == (3, 6)


def _worker_frames(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from videotgb_b200.distributed import all_gather_frame_rows, all_gather_selected_tokens, owned_selection
        g = torch.Generator().manual_seed(1)
        b, t_local, d, nframe = 2, 5, 16, 6
        full = torch.randn(b, world * t_local, d, generator=g)                  # every rank can rebuild the truth
        got = all_gather_frame_rows(full[:, rank * t_local: (rank + 1) * t_local].contiguous())
        ok = torch.equal(got, full)
        # selection spread unevenly over the ranks (4 frames on rank 0, 2 on rank 1), and one rank owning nothing
        for sel in ([0, 1, 3, 4, 7, 9], [5, 5, 6, 7, 8, 9]):
            idx = torch.tensor(sel)
            mine, first, counts = owned_selection(idx, rank, t_local)
            toks = torch.randn(world * t_local, 4, d, generator=torch.Generator().manual_seed(2))   # token of global frame f
            local = toks[[m + rank * t_local for m in mine]] if mine else torch.zeros(0, 4, d)
            out = all_gather_selected_tokens(local, counts, nframe)
            ok = ok and torch.equal(out, toks[sel]) and sum(counts) == nframe and len(mine) == counts[rank] if rank < len(counts) else ok
        q.put((rank, bool(ok)))
    except Exception as e:                                  # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_frame_sharded_gathers_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_frames, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)], res


def test_shard_videos_covers_everything():
    for n in (1, 7, 32, 33):
        for w in (1, 2, 4, 8):
            spans = [shard_videos(n, r, w) for r in range(w)]
            covered = [i for lo, hi in spans for i in range(lo, hi)]
            assert covered == list(range(n))
