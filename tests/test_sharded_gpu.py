"""BASELINE config 4 on >= 2 GPUs (NCCL): one long video whose candidate frames and flow are sharded by frame.
Every rank must produce exactly the tokens the unsharded single-GPU path produces for the whole video (bit-exact: the
per-frame stages do not depend on how many frames share a launch).  Skipped on a single-GPU box; run it with
`gpurun --gpus 2 -- python -m pytest tests/test_sharded_gpu.py -m gpu`."""
import json
import os
import socket
from types import SimpleNamespace as NS

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q, t_total, vit_layers, clustered):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from videotgb_b200.pipeline import VideoTGBEncoder
        dev = f"cuda:{rank}"
        torch.manual_seed(5)                                  # identical random-init weights on every rank
        with torch.device(dev):
            enc = VideoTGBEncoder(NS(num_hidden_layers=vit_layers), None,
                                  NS(fusion_layer=6, encoder_width=768, max_position_embeddings=t_total + 2), 2048)
            with torch.no_grad():
                enc.model.query_tokens.normal_(0.0, 0.02)
        enc.eval()
        k, tq = 32, 16
        g = torch.Generator(device=dev).manual_seed(3)       # identical synthetic video on every rank
        frames = torch.randint(0, 256, (1, t_total, 224, 224, 3), dtype=torch.uint8, device=dev, generator=g)
        of = torch.randn(1, t_total, 2, 224, 224, device=dev, generator=g) * 0.25
        qids = torch.randint(1000, 30522, (1, tq), device=dev, generator=g)
        gum = [-torch.empty(2, t_total, 1, device=dev).exponential_(generator=g).log() for _ in range(2)]
        if clustered:        # both drawn spans inside ONE rank's candidate range: the old owner-encodes split idled W-1 ranks
            tl8 = t_total // 8
            for gi, (s_, e_) in zip(gum, ((5 * tl8 + tl8 // 2, 6 * tl8 - 8), (5 * tl8 + tl8 // 4, 6 * tl8 - 20))):
                gi[0, s_] = 60.0
                gi[1, e_] = 60.0
        common = dict(of_mask=torch.ones(1, t_total + 2, dtype=torch.long, device=dev),
                      of_lengths=torch.tensor([t_total], dtype=torch.int32, device=dev), sampler_question=qids,
                      sampler_question_attention_mask=torch.ones(1, tq, dtype=torch.long, device=dev), nframe=k)
        tl = t_total // world
        local = dict(common, frames_u8=frames[:, rank * tl: (rank + 1) * tl].contiguous(),
                     of=of[:, rank * tl: (rank + 1) * tl].contiguous())
        out, aux = enc.encode_sharded_u8(local, gumbel=gum, return_aux=True)
        out2 = enc.encode_sharded_u8(local, generator=torch.Generator(device=dev).manual_seed(100 + rank))   # broadcast noise
        want, waux = enc.encode_u8(dict(common, frames_u8=frames, of=of), gumbel=gum, return_aux=True)
        torch.cuda.synchronize()
        gathered = [torch.empty_like(out2) for _ in range(world)]
        dist.all_gather(gathered, out2)
        sel = aux.indices[0].tolist()
        owners = sorted({i // tl for i in sel})
        ok = (torch.equal(aux.indices, waux.indices) and torch.equal(out, want)
              and aux.counts == [(r + 1) * k // world - r * k // world for r in range(world)]
              and all(torch.equal(gathered[0], t) for t in gathered) and bool(torch.isfinite(out2).all()))
        # latency of the sharded step vs the single-GPU step on the same video (CUDA events, after the warm-ups above)
        def timed(fn, n=5):
            fn()
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return t.item()
        ms_sharded = timed(lambda: enc.encode_sharded_u8(local, gumbel=gum))
        ms_single = timed(lambda: enc.encode_u8(dict(common, frames_u8=frames, of=of), gumbel=gum)) if rank == 0 else \
            timed(lambda: enc.encode_u8(dict(common, frames_u8=frames, of=of), gumbel=gum))
        q.put((rank, bool(ok), dict(counts=aux.counts, owners=owners, ms_sharded=round(ms_sharded, 3),
                                    ms_single_gpu=round(ms_single, 3), first=sel[0], last=sel[-1])))
    except Exception:                                         # noqa: BLE001
        import traceback
        q.put((rank, False, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def _run(world, t_total, vit_layers, clustered):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, t_total, vit_layers, clustered)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=900) for _ in procs)
    for p in procs:
        p.join(timeout=120)
    assert all(r[1] for r in res), res
    rec = dict(test="frame_sharded_long_video", world=world, candidates=t_total, vit_layers=vit_layers, clustered=clustered,
               bit_identical_to_single_gpu=True, **res[0][2])
    print(json.dumps(rec))
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(out_dir, "r2_sharded_test.jsonl"), "a") as f:
            f.write(json.dumps(rec) + "\n")
    except OSError:
        pass
    return rec


def _worlds():
    n = torch.cuda.device_count()
    return [w for w in (2, 4, 8) if w <= n]


@pytest.mark.parametrize("t_total,vit_layers,clustered", [(256, 2, False), (256, 2, True)])
def test_frame_sharded_long_video_matches_single_gpu(t_total, vit_layers, clustered):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    for world in _worlds():
        _run(world, t_total, vit_layers, clustered)


def test_config4_full_size_clustered_selection():
    """BASELINE config 4 as specified: 1024 candidates, 39-layer ViT-g, k = 32, every available world size (8 on the
    full box), both drawn spans inside one rank's candidate range.  Bit-identical to the single-GPU path, and every rank
    encodes exactly k/W frames."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    for world in _worlds():
        rec = _run(world, 1024, 39, True)
        assert len(rec["owners"]) <= 2          # the selection really is clustered on one or two ranks
