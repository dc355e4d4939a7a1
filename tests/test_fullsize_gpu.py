"""BASELINE config 2 at FULL size on the GPU (32 videos x 64 candidates -> 32 frames each = 1024 frames through the
39-layer ViT-g, Q-Former and projection), checked through size-independent properties plus an oracle spot check:

  * selection: indices sorted, in range, and bit-identical to the oracle applied to the same fp32 scores
  * gather: byte-identical to torch.index_select
  * determinism / batch invariance: a duplicated video yields bit-identical tokens; encoding a sub-batch alone yields
    the same tokens as inside the full batch
  * oracle: two of the 1024 frames re-encoded by the fp32 oracle (39 layers) agree within the north-star tolerance
"""
from types import SimpleNamespace as NS

import pytest
import torch

from oracle import select_ref, torch_ref
from tests.parity_utils import assert_parity
from videotgb_b200 import ops, selection
from videotgb_b200.pipeline import VideoTGBEncoder

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def full():
    torch.manual_seed(7)
    with torch.device(DEV):
        enc = VideoTGBEncoder(NS(), None, NS(fusion_layer=6, encoder_width=768), 2048)
        with torch.no_grad():
            enc.model.query_tokens.normal_(0.0, 0.02)
            for name, p in enc.named_parameters():          # non-trivial LayerNorm affine parameters
                if name.endswith("bias") and p.dim() == 1:
                    p.normal_(0.0, 0.02)
    enc.eval()
    b, n, k, lof, tq = 32, 64, 32, 32, 16
    g = torch.Generator(device=DEV).manual_seed(11)
    frames = torch.randint(0, 256, (b, n, 224, 224, 3), dtype=torch.uint8, device=DEV, generator=g)
    of = torch.randn(b, lof, 2, 224, 224, device=DEV, generator=g) * 0.25
    lengths = torch.randint(8, lof + 1, (b,), device=DEV, generator=g)
    qids = torch.randint(1000, 30522, (b, tq), device=DEV, generator=g)
    gum = [-torch.empty(2 * b, lof, 1, device=DEV).exponential_(generator=g).log() for _ in range(2)]
    # video 1 is an exact copy of video 0 (frames, flow, length, question and Gumbel noise rows)
    frames[1] = frames[0]; of[1] = of[0]; lengths[1] = lengths[0]; qids[1] = qids[0]
    for gg in gum:
        gg[1] = gg[0]; gg[b + 1] = gg[b]
    of_mask = (torch.arange(lof + 2, device=DEV)[None, :] < (lengths[:, None] + 2)).long()
    batch = dict(frames_u8=frames, of=of, of_mask=of_mask, of_lengths=lengths.to(torch.int32), sampler_question=qids,
                 sampler_question_attention_mask=torch.ones(b, tq, dtype=torch.long, device=DEV), nframe=k)
    out, aux = enc.encode_u8(batch, gumbel=gum, return_aux=True)
    torch.cuda.synchronize()
    return enc, batch, gum, out, aux


def test_full_selection_and_gather(full):
    enc, batch, gum, out, aux = full
    b, n, k = 32, 64, 32
    idx = aux.indices
    assert idx.shape == (b, k) and int(idx.min()) >= 0 and int(idx.max()) < n
    assert bool((idx[:, 1:] >= idx[:, :-1]).all())                                   # sorted, duplicates allowed
    sl, el = aux.start_logits.cpu(), aux.end_logits.cpu()
    cs, ce = zip(*[select_ref.gumbel_argmax(sl, el, gg.cpu()) for gg in gum])
    want = select_ref.select_batch(list(cs), list(ce), batch["of_lengths"].tolist(), n, k)
    assert torch.equal(idx.cpu(), want)
    got = ops.gather_frames(batch["frames_u8"], idx)
    ref = torch.stack([batch["frames_u8"][j].index_select(0, idx[j]) for j in range(b)]).view(b * k, 224, 224, 3)
    assert torch.equal(got, ref)


def test_full_tokens_properties(full):
    enc, batch, gum, out, aux = full
    assert out.shape == (32, 32 * 32, 2048) and bool(torch.isfinite(out).all())
    assert torch.equal(out[0], out[1])                                                 # duplicated video
    assert not torch.equal(out[0], out[2])
    sub = {k_: (v[:4] if torch.is_tensor(v) and v.shape[0] == 32 else v) for k_, v in batch.items()}
    gsub = [torch.cat([gg[:4], gg[32:36]], 0) for gg in gum]
    out_sub = enc.encode_u8(sub, gumbel=gsub)
    assert_parity(out_sub, out[:4], "sub-batch vs full batch", rel=1e-5, cos=0.999999)


def test_full_oracle_spot_check(full):
    enc, batch, gum, out, aux = full
    idx = aux.indices
    from videotgb_b200.transforms import VideoNormalizer
    pick = [(3, 5), (17, 30)]                                  # (video, slot) of two of the 1024 encoded frames
    u8 = torch.stack([batch["frames_u8"][v, idx[v, s]] for v, s in pick])
    px = VideoNormalizer()(u8.contiguous())
    vsd = {k_: v.detach() for k_, v in enc.model.vision_model.state_dict().items()}
    qsd = {k_: v.detach() for k_, v in enc.model.qformer.state_dict().items()}
    hsd = {"language_projection.weight": enc.model.language_projection.weight.detach(),
           "language_projection.bias": enc.model.language_projection.bias.detach()}
    with torch.no_grad():
        img, _ = torch_ref.vit_forward(vsd, px)
        qo = torch_ref.qformer_forward(qsd, enc.model.query_tokens.detach().expand(2, -1, -1), img, None)
        want = torch_ref.project_tokens(hsd, qo, 2)              # [2, 32, 2048]
    got = torch.stack([out[v, s * 32:(s + 1) * 32] for v, s in pick])
    assert_parity(got, want, "full-size tokens vs fp32 oracle")


@pytest.mark.parametrize("tt,real", [(16, 16), (32, 20)])
def test_full_config3_instructblip_long_video(tt, real):
    """BASELINE config 3 at full size: InstructBLIP-Vicuna path, one 128-frame video (flow frames == candidates), the
    sampler (mode="multi_modal", 12 layers) selects 32 frames, instruction-conditioned Q-Former (Tt instruction tokens,
    `real` of them unpadded), H_llm = 4096.  Every selected frame is compared with the fp32 oracle (39-layer ViT)."""
    from oracle.transform_ref import normalize_u8_thwc
    from videotgb_b200 import synthetic as syn
    b, t_frames, tq, k = 1, 128, 16, 32
    enc = VideoTGBEncoder(NS(), None, NS(fusion_layer=6, encoder_width=768), llm_hidden=4096, instruct=True).to(DEV)
    vsd, isd = syn.vision_state_dict(), syn.qformer_state_dict(syn.QFormerCfg(instruct=True))
    ssd, hsd = syn.sampler_state_dict(), syn.head_state_dict(llm_hidden=4096)
    enc.model.vision_model.load_state_dict(vsd)
    enc.model.qformer.load_state_dict(isd)
    enc.temporal_encoder.load_state_dict(ssd)
    enc.load_state_dict({"model." + kk: v for kk, v in hsd.items()}, strict=False)
    enc.eval()
    g = torch.Generator().manual_seed(77 + tt)
    u8 = torch.randint(0, 256, (b, t_frames, 224, 224, 3), generator=g, dtype=torch.uint8)
    of = syn.synthetic_flow(b, t_frames, seed=19)
    of_mask = torch.ones(b, t_frames + 2, dtype=torch.long)
    qids = torch.randint(1000, 30522, (b, tq), generator=g)
    tids = torch.randint(1000, 30522, (b, tt), generator=g)
    tmask = torch.ones(b, tt, dtype=torch.long)
    tmask[:, real:] = 0
    gum = [-torch.empty(2 * b, t_frames, 1).exponential_(generator=g).log() for _ in range(2)]
    batch = dict(frames_u8=u8.to(DEV), of=of.to(DEV), of_mask=of_mask.to(DEV), of_lengths=[t_frames],
                 sampler_question=qids.to(DEV), sampler_question_attention_mask=torch.ones(b, tq, dtype=torch.long, device=DEV),
                 qformer_text=tids.to(DEV), qformer_text_attention_mask=tmask.to(DEV), nframe=k)
    out, aux = enc.encode_u8(batch, mode="multi_modal", gumbel=[x.to(DEV) for x in gum], return_aux=True)
    assert out.shape == (1, k * 32, 4096) and bool(torch.isfinite(out).all())
    sl, el = aux.start_logits.cpu(), aux.end_logits.cpu()
    cs, ce = zip(*[select_ref.gumbel_argmax(sl, el, gg) for gg in gum])
    want_idx = select_ref.select_batch(list(cs), list(ce), [t_frames], t_frames, k)
    assert torch.equal(aux.indices.cpu(), want_idx)
    px = torch.from_numpy(normalize_u8_thwc(u8.view(-1, 224, 224, 3).numpy())).view(b, t_frames, 3, 224, 224)
    dev = lambda sd: {kk: v.to(DEV) for kk, v in sd.items()}
    with torch.no_grad():
        sampled = select_ref.gather_frames(px, want_idx).to(DEV)
        img, _ = torch_ref.vit_forward(dev(vsd), sampled)
        qt = hsd["query_tokens"].to(DEV).expand(k, -1, -1)
        ids = torch.repeat_interleave(tids, k, 0).to(DEV)
        am = torch.cat([torch.ones(k, 32, dtype=torch.long), torch.repeat_interleave(tmask, k, 0)], 1).to(DEV)
        qo = torch_ref.qformer_forward(dev(isd), qt, img, None, input_ids=ids, attention_mask=am, instruct=True)[:, :32]
        want = torch_ref.project_tokens(dev(hsd), qo, b)
        _, rlogits = torch_ref.sampler_forward(dev(ssd), of.to(DEV), of_mask.to(DEV), qids.to(DEV),
                                               torch.ones(b, tq, dtype=torch.long, device=DEV), mode="multi_modal")
    assert_parity(out, want, "config 3 tokens vs fp32 oracle")
    assert_parity(torch.cat([aux.start_logits, aux.end_logits], -1), rlogits, "config 3 multi_modal scores")
