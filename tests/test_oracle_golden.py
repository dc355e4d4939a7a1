"""CPU: the oracle restatement vs. fixtures produced by RUNNING the reference (oracle/gen_golden.py)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import cases, raft_ref, select_ref, torch_ref, transform_ref
from videotgb_b200 import synthetic as syn

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _close(t, meta, arr, stride, tol=2e-5):
    t = t.detach().float()
    assert list(t.shape) == meta["shape"]
    ref = torch.from_numpy(arr)
    got = t.flatten()[::stride]
    err = (got - ref).abs().max().item()
    assert err <= tol * max(1.0, meta["absmax"]), (err, meta)
    assert abs(float(t.std()) - meta["std"]) < 1e-3 * max(1.0, meta["std"])


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(GOLDEN, "models_meta.json")) as f:
        meta = json.load(f)
    return meta, np.load(os.path.join(GOLDEN, "models.npz"))


def test_select_kat_bit_exact():
    with open(os.path.join(GOLDEN, "select_kat.json")) as f:
        kat = json.load(f)
    assert len(kat) >= 100
    for c in kat:
        kw = dict(eval=dict(mapping="scaled", inclusive=False), train=dict(mapping="scaled", inclusive=True),
                  ratio=dict(mapping="ratio", inclusive=False))[c["variant"]]
        for j, length in enumerate(c["of_lengths"]):
            starts = [d[0] for d in c["draws"][j]]
            ends = [d[1] for d in c["draws"][j]]
            got = select_ref.spans_to_indices(starts, ends, length, c["num_frames"], c["nframe"], **kw)
            assert got == c["expect"][j], (c["name"], j, got, c["expect"][j])


def test_survey_kat_table():
    # SURVEY.md section 8c known-answer table (derived there by executing the reference statements)
    f = select_ref.spans_to_indices
    assert f([0, 0], [0, 0], 64, 32, 4) == [3, 10, 18, 26]
    assert f([0, 0], [0, 0], 64, 32, 4, inclusive=True) == [3, 11, 19, 27]
    assert f([10, 40], [30, 50], 64, 32, 4) == [5, 8, 12, 21]
    assert f([10, 40], [30, 50], 64, 32, 4, inclusive=True) == [5, 9, 13, 22]
    assert f([30, 5], [10, 6], 64, 32, 4) == [3, 11, 19, 27]
    assert f([20, 20], [22, 22], 64, 32, 4) == [9, 9, 9, 9]
    assert f([70, 3], [10, 9], 64, 32, 4) == [3, 10, 18, 26]
    assert f([100, 350], [400, 700], 1024, 1024, 32)[:4] == [108, 127, 146, 165]
    assert f([5, 50], [100, 120], 128, 32, 32, mapping="ratio")[:6] == [1, 1, 2, 3, 4, 5]
    # (SURVEY's 'uniform baseline' row is off by one from index 1 on; executing the reference expression gives:)
    assert select_ref.uniform_indices(32, 8) == [1, 5, 9, 13, 17, 21, 25, 29]


def _pseudo_kat():
    with open(os.path.join(GOLDEN, "pseudo_span_kat.json")) as f:
        return json.load(f)


def test_pseudo_span_and_span_ce_vs_executed_reference():
    kat = _pseudo_kat()
    assert len(kat["span"]) >= 50
    for c in kat["span"]:
        for j, length in enumerate(c["of_lengths"]):
            # the reference works on `tensor.tolist()` of a float32 tensor
            row = torch.tensor(c["scores"][j], dtype=torch.float32).tolist()
            assert select_ref.pseudo_span(row, length) == (c["start"][j], c["end"][j])
    for c in kat["ce"]:
        lg = torch.tensor(c["logits"], dtype=torch.float32)
        s, e = lg.split(1, dim=-1)
        got = float(select_ref.span_ce_loss(s, e, torch.tensor(c["start"]), torch.tensor(c["end"])))
        assert (np.isnan(got) and np.isnan(c["loss"])) or abs(got - c["loss"]) <= 1e-6 * max(1.0, abs(c["loss"]))


def test_rouge_n_host_mirror_vs_executed_reference():
    from videotgb_b200.selection import rouge_n
    for c in _pseudo_kat()["rouge"]:
        assert rouge_n(c["gold"], c["pred"], c["ignore"]) == c["score"], c


def test_gumbel_argmax_matches_reference_draw():
    z = np.load(os.path.join(GOLDEN, "gumbel.npz"))
    for tag in "abc":
        s, e, g = (torch.from_numpy(z[f"{tag}_{k}"]) for k in ("start", "end", "gumbel"))
        si, ei = select_ref.gumbel_argmax(s, e, g)
        idx = torch.from_numpy(z[f"{tag}_index"])[:, 0]
        assert torch.equal(torch.cat([si, ei]), idx)


def test_transform_bit_exact():
    z = np.load(os.path.join(GOLDEN, "transform.npz"))
    got = transform_ref.normalize_u8_thwc(z["u8"])
    assert got.dtype == np.float32 and np.array_equal(got, z["norm"])
    assert np.array_equal(transform_ref.resize_to_uint8(z["clip"], 16), z["clip_u8"])


def test_resize_restatement_vs_reference_chain_and_live_interpolate():
    z = np.load(os.path.join(GOLDEN, "resize.npz"))
    for tag in ("down", "up", "mixed", "same"):
        got = transform_ref.resize_bilinear_u8(z[f"{tag}_in"], 32)
        assert np.array_equal(got, z[f"{tag}_u8"]), tag
        assert np.array_equal(transform_ref.normalize_u8_thwc(got), z[f"{tag}_norm"]), tag
    # the ATen arithmetic the restatement encodes, checked live at the sizes real videos come in
    rng = np.random.default_rng(3)
    for h, w in [(360, 640), (240, 320), (480, 854), (225, 223), (2, 2)]:
        u8 = rng.integers(0, 256, size=(2, h, w, 3), dtype=np.uint8)
        u8[0] = 200
        clip = torch.from_numpy(u8).permute(3, 0, 1, 2).float()
        want = torch.nn.functional.interpolate(clip, size=(224, 224), mode="bilinear", align_corners=False)
        want = want.to(torch.uint8).permute(1, 2, 3, 0).numpy()
        assert np.array_equal(transform_ref.resize_bilinear_u8(u8, 224), want), (h, w)


def test_raft_restatement_vs_reference_class():
    z = np.load(os.path.join(GOLDEN, "raft.npz"))
    sd = syn.raft_state_dict()
    img1, img2 = cases.raft_pair()
    with torch.no_grad():
        flow, aux = raft_ref.raft_forward(sd, img1, img2, iters=20, return_aux=True)
    assert abs(float(flow.abs().max()) - float(z["flow_absmax"])) <= 1e-3 * float(z["flow_absmax"])
    assert np.abs(flow[:, :, ::4, ::4].numpy() - z["flow"]).max() <= 2e-3 * float(z["flow_absmax"])     # fp32 reassociation only
    assert np.abs(aux["fmap1"][:, ::8, ::2, ::2].numpy() - z["fmap1"]).max() <= 1e-4 * np.abs(z["fmap1"]).max()


def test_vit2_and_heads(gold):
    meta, arr = gold
    cfg = syn.VisionCfg(layers=2)
    sd = syn.vision_state_dict(cfg)
    with torch.no_grad():
        last, pool = torch_ref.vit_forward(sd, cases.vit_pixels(2), layers=2)
        _close(last, meta["vit2_last"], arr["vit2_last"], 7)
        _close(pool, meta["vit2_pool"], arr["vit2_pool"], 1)
        hsd = syn.head_state_dict()
        qt = hsd["query_tokens"].expand(2, -1, -1)
        qo = torch_ref.qformer_forward(syn.qformer_state_dict(), qt, last,
                                       torch.ones(last.shape[:2], dtype=torch.long))
        _close(qo, meta["qf_out"], arr["qf_out"], 3)
        proj = torch_ref.project_tokens(hsd, qo, 1)
        _close(proj, meta["qf_proj"], arr["qf_proj"], 5)
        ids, amask, _ = cases.instruct_text(2)
        assert np.array_equal(ids.numpy(), arr["iq_ids"]) and np.array_equal(amask.numpy(), arr["iq_mask"])
        io = torch_ref.qformer_forward(syn.qformer_state_dict(syn.QFormerCfg(instruct=True)), qt, last,
                                       torch.ones(last.shape[:2], dtype=torch.long), input_ids=ids,
                                       attention_mask=amask, instruct=True)
        _close(io, meta["iq_out"], arr["iq_out"], 3)
        # mean-pool assembly == reference expression (eval/utils/model.py:186-195)
        mp = torch_ref.project_tokens(hsd, qo, 1, mode="mean")
        ref = torch.nn.functional.linear(qo.mean(0, keepdim=True), hsd["language_projection.weight"],
                                         hsd["language_projection.bias"])
        assert torch.allclose(mp, ref, atol=1e-6)


def test_vit39(gold):
    meta, arr = gold
    sd = syn.vision_state_dict(syn.VisionCfg())
    with torch.no_grad():
        last, pool = torch_ref.vit_forward(sd, cases.vit_pixels(1))
    _close(last, meta["vit39_last"], arr["vit39_last"], 7, tol=5e-5)


def test_sampler(gold):
    meta, arr = gold
    sd = syn.sampler_state_dict()
    of, of_mask, of_len, qids, qmask = cases.sampler_inputs()
    assert np.array_equal(qids.numpy(), arr["s_qids"]) and np.array_equal(of_mask.numpy(), arr["s_ofmask"])
    with torch.no_grad():
        emb = torch_ref.of_embedding(sd, of, of_mask)
        _close(emb, meta["s_ofembed"], arr["s_ofembed"], 1)
        emb2 = torch_ref.of_embedding_reassociated(sd, of, of_mask)
        assert (emb - emb2).abs().max().item() < 2e-5
        for mode in ("fusion", "multi_modal"):
            seq, logits = torch_ref.sampler_forward(sd, of, of_mask, qids, qmask, mode=mode)
            _close(seq, meta[f"s_{mode}_seq"], arr[f"s_{mode}_seq"], 3)
            _close(logits, meta[f"s_{mode}_logits"], arr[f"s_{mode}_logits"], 1)
