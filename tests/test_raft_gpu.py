"""the RAFT optical-flow extractor on the sm_100a kernels vs the reference (golden fixture from
the reference class) and vs the fp32 oracle run on the device; plus the individual flow kernels."""
import os

import numpy as np
import pytest
import torch

from oracle import cases, raft_ref
from tests.parity_utils import assert_parity
from videotgb_b200 import ops
from videotgb_b200 import synthetic as syn
from videotgb_b200.raft import RAFT, InputPadder

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda"


@pytest.fixture(scope="module")
def raft():
    m = RAFT().to(DEV)
    m.load_state_dict(syn.raft_state_dict())
    return m.eval()


def test_im2col_matches_unfold():
    g = torch.Generator().manual_seed(1)
    for n, h, w, c, kh, kw, st, pad in [(2, 12, 10, 16, 3, 3, (1, 1), (1, 1)), (1, 9, 11, 8, 1, 5, (1, 1), (0, 2)),
                                        (2, 14, 14, 24, 3, 3, (2, 2), (1, 1)), (1, 16, 16, 3, 7, 7, (2, 2), (3, 3)),
                                        (1, 8, 8, 2, 7, 7, (1, 1), (3, 3)), (2, 10, 10, 32, 1, 1, (2, 2), (0, 0))]:
        x = torch.randn(n, c, h, w, generator=g).to(torch.bfloat16)
        cols = torch.nn.functional.unfold(x.float(), (kh, kw), padding=pad, stride=st)            # [n, c*kh*kw, L], (c,ky,kx)
        want = cols.view(n, c, kh * kw, -1).permute(0, 3, 2, 1).reshape(-1, kh * kw * c)          # rows (n,oy,ox), cols (tap,c)
        cpad = (c + 7) // 8 * 8
        rows = torch.zeros(n * h * w, cpad, dtype=torch.bfloat16)
        rows[:, :c] = x.permute(0, 2, 3, 1).reshape(-1, c)
        got = ops.im2col_nhwc(rows.to(DEV), n, h, w, c, kh, kw, st, pad).cpu().float()
        assert torch.equal(got[:, : kh * kw * c], want) and got[:, kh * kw * c:].abs().sum() == 0, (n, h, w, c, kh, kw)


def test_instance_norm_and_gates():
    g = torch.Generator().manual_seed(2)
    x = torch.randn(3, 35, 40, generator=g) * 2 + 0.7                                           # [n, hw, c]
    want = torch.nn.functional.instance_norm(x.permute(0, 2, 1).reshape(3, 40, 5, 7)).reshape(3, 40, 35).permute(0, 2, 1)
    of, ob = ops.instance_norm_nhwc(x.reshape(-1, 40).to(DEV), 3, 35, True, True, True)
    assert (of.cpu() - want.relu().reshape(-1, 40)).abs().max() < 2e-5
    assert (ob.cpu().float() - want.relu().reshape(-1, 40)).abs().max() < 2e-2
    m, c = 50, 128
    zr, q, h = torch.randn(m, 2 * c, generator=g), torch.randn(m, c, generator=g), torch.randn(m, c, generator=g)
    rh = torch.zeros(m, 384, dtype=torch.bfloat16, device=DEV)
    ops.gru_rh(zr.to(DEV), h.to(DEV), rh[:, :c])
    assert (rh[:, :c].cpu().float() - torch.sigmoid(zr[:, c:]) * h).abs().max() < 2e-2
    hd = h.to(DEV).clone()
    hb = torch.zeros(m, 384, dtype=torch.bfloat16, device=DEV)
    ops.gru_update(zr.to(DEV), q.to(DEV), hd, hb[:, :c])
    z = torch.sigmoid(zr[:, :c])
    assert (hd.cpu() - ((1 - z) * h + z * torch.tanh(q))).abs().max() < 1e-5


def test_corr_pyramid_lookup_and_upsample():
    g = torch.Generator().manual_seed(3)
    n, h, w = 2, 16, 20
    f1, f2 = torch.randn(n, 256, h, w, generator=g), torch.randn(n, 256, h, w, generator=g)
    pyr = raft_ref.corr_pyramid(f1, f2)
    coords0 = raft_ref.coords_grid(n, h, w, "cpu")
    flow = torch.randn(n, 2, h, w, generator=g) * 3
    want = raft_ref.corr_lookup(pyr, coords0 + flow)                                            # [n, 324, h, w]
    dpyr = [pyr[0].view(-1, h, w).to(DEV)]
    for _ in range(3):
        dpyr.append(ops.corr_pool(dpyr[-1]))
    for a, b in zip(dpyr, pyr):
        assert (a.cpu() - b.view(a.shape)).abs().max() < 1e-5
    out = torch.zeros(n * h * w, 328, dtype=torch.bfloat16, device=DEV)
    ops.corr_lookup(dpyr, flow.permute(0, 2, 3, 1).reshape(-1, 2).contiguous().to(DEV), n, h, w, out)
    got = out[:, :324].cpu().float().view(n, h, w, 324).permute(0, 3, 1, 2)
    assert (got - want).abs().max() <= 2e-2 * want.abs().max() and out[:, 324:].abs().max() == 0
    mask = torch.randn(n, 576, h, w, generator=g)
    want_up = raft_ref.upsample_flow(flow, mask)
    got_up = ops.raft_upsample(mask.permute(0, 2, 3, 1).reshape(-1, 576).contiguous().to(DEV),
                               flow.permute(0, 2, 3, 1).reshape(-1, 2).contiguous().to(DEV), n, h, w).cpu()
    assert (got_up - want_up).abs().max() <= 1e-4 * want_up.abs().max()


def test_raft_matches_reference_fixture_and_oracle(raft):
    z = np.load(os.path.join(GOLDEN, "raft.npz"))
    img1, img2 = cases.raft_pair()
    flow = raft(img1.to(DEV), img2.to(DEV))
    assert flow.shape == (2, 2, 128, 160)
    sd = {k: v.to(DEV) for k, v in syn.raft_state_dict().items()}
    with torch.no_grad():
        want = raft_ref.raft_forward(sd, img1.to(DEV), img2.to(DEV), iters=20)
    assert_parity(flow, want, "raft flow vs fp32 oracle (device)")
    assert_parity(flow[:, :, ::4, ::4].cpu(), torch.from_numpy(z["flow"]), "raft flow vs reference fixture")


def test_raft_224_pairs_like_the_offline_pipeline(raft):
    """eval/utils/model.py:77-84: pad, of_extractor(frames[:-1], frames[1:]), repeat the last flow."""
    g = torch.Generator().manual_seed(5)
    base = torch.nn.functional.interpolate(torch.rand(1, 3, 30, 30, generator=g), size=(224 + 8, 224 + 8), mode="bicubic",
                                           align_corners=False).clamp(0, 1) * 255.0
    frames = torch.cat([base[:, :, i: i + 224, 2 * i // 2: 2 * i // 2 + 224] for i in range(4)], 0).contiguous()   # a slow pan
    padder = InputPadder(frames[0].shape)
    fr = padder.pad(frames.to(DEV))
    flow = raft(fr[:-1], fr[1:])
    assert flow.shape == (3, 2, 224, 224) and bool(torch.isfinite(flow).all())
    sd = {k: v.to(DEV) for k, v in syn.raft_state_dict().items()}
    with torch.no_grad():
        want = raft_ref.raft_forward(sd, fr[:-1], fr[1:], iters=20)
    assert_parity(flow, want, "raft 224 flow vs fp32 oracle (device)")


def test_extract_flow_feeds_the_sampler(raft):
    """eval/utils/model.py:75-100: RAFT flow of each video -> `of` [B,T,2,H,W] -> temporal_encoder logits."""
    from types import SimpleNamespace as NS
    from videotgb_b200.pipeline import VideoTGBEncoder
    with torch.device(DEV):
        enc = VideoTGBEncoder(NS(num_hidden_layers=1), None, NS(fusion_layer=6, encoder_width=768), 2048, with_flow_extractor=True)
    enc.of_extractor.load_state_dict(syn.raft_state_dict())
    enc.eval()
    g = torch.Generator().manual_seed(6)
    frames = torch.rand(2, 5, 3, 224, 224, generator=g).to(DEV) * 255
    of = enc.extract_flow(frames)
    assert of.shape == (2, 5, 2, 224, 224) and torch.equal(of[:, -1], of[:, -2])
    want = raft(frames[0, :-1], frames[0, 1:])
    assert torch.equal(of[0, :-1], want)
    qids = torch.randint(1000, 30522, (2, 8), generator=g).to(DEV)
    _, logits = enc.temporal_encoder(encoder_embeds=of, attention_mask=torch.ones(2, 7, dtype=torch.long, device=DEV),
                                     encoder_hidden_states=qids, mode="multi_modal")
    assert logits.shape == (2, 5, 2) and bool(torch.isfinite(logits).all())
