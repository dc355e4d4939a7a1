"""BASELINE config 4 on >= 2 GPUs (NCCL): one long video whose candidate frames and flow are sharded by frame.
Every rank must produce exactly the tokens the unsharded single-GPU path produces for the whole video (bit-exact: the
per-frame stages do not depend on how many frames share a launch).  Skipped on a single-GPU box; run it with
`gpurun --gpus 2 -- python -m pytest tests/test_sharded_gpu.py -m gpu`."""
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


def _worker(rank, world, port, q, t_total, vit_layers):
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
        ok = (torch.equal(aux.indices, waux.indices) and torch.equal(out, want) and sum(aux.counts) == k
              and all(torch.equal(gathered[0], t) for t in gathered) and bool(torch.isfinite(out2).all()))
        q.put((rank, bool(ok), aux.counts))
    except Exception as e:                                    # noqa: BLE001
        q.put((rank, False, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("t_total,vit_layers", [(256, 2)])
def test_frame_sharded_long_video_matches_single_gpu(t_total, vit_layers):
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if world < 4 else 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, t_total, vit_layers)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=120)
    assert all(r[1] for r in res), res
