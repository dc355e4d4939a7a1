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
        from videotgb_b200.distributed import (all_gather_encoded_tokens, all_gather_frame_rows, encode_plan,
                                               redistribute_selected_frames)
        g = torch.Generator().manual_seed(1)
        b, t_local, d = 2, 5, 16
        full = torch.randn(b, world * t_local, d, generator=g)                  # every rank can rebuild the truth
        got = all_gather_frame_rows(full[:, rank * t_local: (rank + 1) * t_local].contiguous())
        ok = torch.equal(got, full)
        # the whole video's u8 frames (every rank can rebuild the truth); rank r holds [r*t_local, (r+1)*t_local)
        video = torch.randint(0, 256, (1, world * t_local, 4, 4, 3), generator=torch.Generator().manual_seed(7), dtype=torch.uint8)
        shard = video[:, rank * t_local: (rank + 1) * t_local].contiguous()
        toks = torch.randn(world * t_local, 4, d, generator=torch.Generator().manual_seed(2)).to(torch.bfloat16)
        # selections: spread over both ranks; clustered entirely in rank 1's range; duplicates; odd k (uneven split)
        for sel in ([0, 1, 3, 4, 7, 9], [5, 5, 6, 7, 8, 9], [0, 0, 1, 1, 2, 2], [2, 3, 8], [1, 2, 3, 4, 6]):
            plan = encode_plan(sel, rank, world, t_local)
            k = len(sel)
            ok = ok and plan.counts == [(r + 1) * k // world - r * k // world for r in range(world)]
            ok = ok and sum(plan.send_splits) == len(plan.send_local) and sum(plan.recv_splits) == plan.hi - plan.lo
            mine = redistribute_selected_frames(shard, plan)
            ok = ok and torch.equal(mine, video[0, sel[plan.lo: plan.hi]])       # the frames this rank ENCODES, in order
            out = all_gather_encoded_tokens(toks[sel[plan.lo: plan.hi]].contiguous(), plan.counts)
            ok = ok and torch.equal(out, toks[sel])
        q.put((rank, bool(ok)))
    except Exception as e:                                  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
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


def test_encode_plan_is_even_and_consistent():
    """Host logic of the config-4 split for every world size: positions are split evenly, and what rank s sends to
    rank d is exactly what d expects from s."""
    from videotgb_b200.distributed import encode_plan
    import random
    rnd = random.Random(0)
    for world in (1, 2, 4, 8):
        t_local = 1024 // world
        for k in (4, 32, 33):
            for _ in range(10):
                lo = rnd.randrange(0, 1024 - 64)
                sel = sorted(rnd.choice(range(lo, lo + 64)) for _ in range(k))          # clustered span, duplicates
                plans = [encode_plan(sel, r, world, t_local) for r in range(world)]
                assert [p.hi - p.lo for p in plans] == plans[0].counts and sum(plans[0].counts) == k
                assert max(plans[0].counts) - min(plans[0].counts) <= 1
                for s_ in range(world):
                    for d_ in range(world):
                        assert plans[s_].send_splits[d_] == plans[d_].recv_splits[s_]
                    assert [i + s_ * t_local for i in plans[s_].send_local] == [i for i in sel if i // t_local == s_]


def test_shard_videos_covers_everything():
    for n in (1, 7, 32, 33):
        for w in (1, 2, 4, 8):
            spans = [shard_videos(n, r, w) for r in range(w)]
            covered = [i for lo, hi in spans for i in range(lo, hi)]
            assert covered == list(range(n))
