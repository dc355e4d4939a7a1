"""Generate tests/golden/* by RUNNING THE REFERENCE in the build container.

    python -m oracle.gen_golden            (needs /root/reference; never runs on the GPU box)

What it pins:
  select_kat.json      -- A8: the reference's inline selection statements are read from
                          /root/reference/src/models/LSTP_SF_blip2_module.py (eval :431-459, train :282-312),
                          LSTP_module.py (ratio mapping) and executed unmodified on real 0-dim int64 tensors.
  pseudo_span_kat.json -- N2: the monotone-stack pseudo-label statements (:193-224), the span cross-entropy
                          statements (:241-253) and rouge_n (my_metrics.py:131-179), executed from the reference source.
  gumbel.npz           -- A7: F.gumbel_softmax(...).argmax with a fixed torch seed, plus the explicit noise.
  resize.npz           -- N3: the reference ResizeVideo/ToUint8/ToTHWC(/ToTensorVideo/NormalizeVideo) chain on decoded frames.
  raft.npz             -- N1: the reference RAFT class on a seeded frame pair (flow_up subsampled, fnet feature sample).
  transform.npz        -- A0: reference ToTensorVideo/NormalizeVideo on random u8 frames.
  task_module.npz      -- (b): the unmodified reference `eval_forward` methods (BLIP-2 and InstructBLIP task modules)
                          on the reference containers built from HF configs: what they pass to `language_model.generate`.
  vit_small.npz ...    -- A10-A16, A1-A6: the reference classes (through reference_shim) loaded with the
                          seeded synthetic state dicts, fp32, small inputs.  Outputs stored as fp16-safe fp32
                          slices + full-tensor statistics to keep fixtures small.
"""
from __future__ import annotations

import json
import os
import sys
import textwrap

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import reference_shim                       # noqa: E402
from videotgb_b200 import synthetic as syn                # noqa: E402

REF = reference_shim.REFERENCE_ROOT


def _ref_lines(rel, lo, hi):
    with open(os.path.join(REF, rel)) as f:
        lines = f.readlines()[lo - 1: hi]
    return textwrap.dedent("".join(lines))


def _run_reference_selection(src, cand_start_index, cand_end_index, of_lengths, num_frames, nframe):
    """Execute the reference's `for j in range(len(video_lengths))` loop body source `src`."""
    b = len(of_lengths)
    pixel_values = torch.arange(b * num_frames, dtype=torch.float32).view(b, num_frames, 1, 1, 1)
    sampled = torch.zeros(b, nframe, 1, 1, 1)
    env = dict(torch=torch, np=np, batch={"of_lengths": list(of_lengths)}, cand_start_index=cand_start_index,
               cand_end_index=cand_end_index, num_frames=num_frames, nframe=nframe, pixel_values=pixel_values,
               sampled_pixel_values=sampled, video_lengths=list(of_lengths))
    exec(src, env)
    idx = (env["sampled_pixel_values"].view(b, nframe) - torch.arange(b).view(b, 1) * num_frames).long()
    return idx.tolist()


def gen_select():
    eval_src = _ref_lines("src/models/LSTP_SF_blip2_module.py", 433, 461)
    train_src = _ref_lines("src/models/LSTP_SF_blip2_module.py", 284, 314)
    ratio_src = _ref_lines("src/models/LSTP_module.py", 233, 263)
    assert "range(start, end)" in eval_src and "range(start, (end+1))" in train_src
    assert "cand_start/video_length*num_frames" in ratio_src
    # (the ratio variant hard-codes video_length = num_frames + 2)
    cases = []
    rng = np.random.default_rng(20240917)

    def add(name, src, variant, draws, lengths, n, k):
        cs = [torch.tensor([d[i][0] for d in draws]) for i in range(len(draws[0]))]
        ce = [torch.tensor([d[i][1] for d in draws]) for i in range(len(draws[0]))]
        out = _run_reference_selection(src, cs, ce, lengths, n, k)
        cases.append(dict(name=name, variant=variant, draws=draws, of_lengths=list(map(int, lengths)),
                          num_frames=n, nframe=k, expect=out))

    # hand-written corner cases, one video each, two draws
    hand = [
        ("fallback", [[(0, 0), (0, 0)]], [64], 32, 4),
        ("two_spans", [[(10, 30), (40, 50)]], [64], 32, 4),
        ("reversed_degenerate", [[(30, 10), (5, 6)]], [64], 32, 4),
        ("tiny_doubling", [[(20, 22), (20, 22)]], [64], 32, 4),
        ("out_of_range", [[(70, 10), (3, 9)]], [64], 32, 4),
        ("long_video", [[(100, 400), (350, 700)]], [1024], 1024, 32),
        ("exact_k", [[(0, 9), (0, 9)]], [32], 32, 4),
        ("len2", [[(0, 1), (1, 1)]], [2], 32, 4),
    ]
    for name, draws, lengths, n, k in hand:
        add(name + "_eval", eval_src, "eval", draws, lengths, n, k)
        add(name + "_train", train_src, "train", draws, lengths, n, k)
    # random batches incl. padded positions being drawn (index >= of_length)
    for t in range(40):
        b = int(rng.integers(1, 9))
        lof = int(rng.choice([8, 32, 64, 128, 510, 1024]))
        n = int(rng.choice([8, 32, 64, lof]))
        k = int(rng.choice([1, 4, 8, 32]))
        lengths = [int(x) for x in rng.integers(2, lof + 1, size=b)]
        draws = [[(int(rng.integers(0, lof)), int(rng.integers(0, lof))) for _ in range(2)] for _ in range(b)]
        add(f"rand{t}_eval", eval_src, "eval", draws, lengths, n, k)
        add(f"rand{t}_train", train_src, "train", draws, lengths, n, k)
    for t in range(20):
        b = int(rng.integers(1, 5))
        n = int(rng.choice([8, 32, 64]))
        k = int(rng.choice([4, 8, 32]))
        draws = [[(int(rng.integers(0, n + 2)), int(rng.integers(0, n + 2))) for _ in range(2)] for _ in range(b)]
        add(f"rand{t}_ratio", ratio_src, "ratio", draws, [n + 2] * b, n, k)
    with open(os.path.join(GOLD, "select_kat.json"), "w") as f:
        json.dump(cases, f)
    print("select cases:", len(cases))


def gen_pseudo_span():
    """N2: execute the reference's monotone-stack pseudo-label statements (LSTP_SF_blip2_module.py:193-224) verbatim."""
    src = _ref_lines("src/models/LSTP_SF_blip2_module.py", 193, 224)
    assert "monotone stack" in src and "flow_lengths" in src
    rng = np.random.default_rng(77)
    cases = []
    for t in range(60):
        b = int(rng.integers(1, 6))
        n = int(rng.choice([4, 8, 32, 64]))
        kind = t % 4
        if kind == 0:
            sc = rng.random((b, n)).astype(np.float32)
        elif kind == 1:
            sc = np.round(rng.random((b, n)) * 4).astype(np.float32) / 4          # many ties
        elif kind == 2:
            sc = np.zeros((b, n), np.float32); sc[:, rng.integers(0, n)] = 1.0      # single peak
        else:
            sc = np.sort(rng.random((b, n)).astype(np.float32), axis=1)           # monotone
        lengths = [int(x) for x in rng.integers(2, 65, size=b)]
        env = dict(torch=torch, scores=torch.tensor(sc).view(-1), batch_size=b, num_frames=n,
                   batch={"of_lengths": lengths}, pixel_values=torch.zeros(1))
        exec(src, env)
        cases.append(dict(scores=sc.tolist(), of_lengths=lengths, start=[int(v) for v in env["start_targets"]],
                          end=[int(v) for v in env["end_targets"]]))
    # span cross-entropy: the reference statements :241-253 executed on random logits (incl. ignored / clamped targets)
    ce_src = _ref_lines("src/models/LSTP_SF_blip2_module.py", 241, 253)
    assert "ignore_index=ignored_index" in ce_src and "mrc_loss" in ce_src
    ce = []
    for t in range(12):
        b, l = int(rng.integers(1, 9)), int(rng.choice([4, 32, 64, 100]))
        logits = torch.tensor(rng.normal(size=(b, l, 2)).astype(np.float32) * 3)
        st = torch.tensor(rng.integers(-2, l + 3, size=b))
        en = torch.tensor(rng.integers(0, l, size=b))
        if t == 0:
            st = torch.full((b,), l)                       # every start row ignored -> NaN
        env = dict(torch=torch, CrossEntropyLoss=torch.nn.CrossEntropyLoss, of_logits=logits, start_targets=st.clone(),
                   end_targets=en.clone())
        exec(ce_src, env)
        ce.append(dict(logits=logits.tolist(), start=st.tolist(), end=en.tolist(), loss=float(env["mrc_loss"])))
    # rouge_n (src/gadgets/my_metrics.py:131-179), executed from the reference source (torchmetrics is not installed)
    env = {}
    exec(_ref_lines("src/gadgets/my_metrics.py", 131, 179), env)
    words = "a the cat sat on mat , . dog runs fast in park it is".split()
    rouge = []
    for t in range(40):
        n = int(rng.integers(1, 5))
        gold = [" ".join(rng.choice(words, size=int(rng.integers(1, 8)))) for _ in range(n)]
        pred = [" ".join(rng.choice(words, size=int(rng.integers(0, 8)))) for _ in range(n)]
        ignore = None if t % 7 == 6 else [",", "."]
        if t % 5 == 4:
            gold, pred = gold[0], pred[0]
        rouge.append(dict(gold=gold, pred=pred, ignore=ignore, score=env["rouge_n"](gold, pred, ignore)))
    with open(os.path.join(GOLD, "pseudo_span_kat.json"), "w") as f:
        json.dump(dict(span=cases, ce=ce, rouge=rouge), f)
    print("pseudo-span cases:", len(cases), "ce:", len(ce), "rouge:", len(rouge))


def gen_gumbel():
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(7)
    out = {}
    for tag, (b, l) in {"a": (4, 32), "b": (3, 130), "c": (2, 1024)}.items():
        start = torch.randn(b, l, 1, generator=g)
        end = torch.randn(b, l, 1, generator=g)
        logits = torch.cat([start, end], dim=0)
        torch.manual_seed(100 + l)
        state = torch.get_rng_state()
        prob = F.gumbel_softmax(logits, tau=0.5, dim=1)                 # the reference statement (:425)
        index = prob.argmax(dim=1)
        torch.set_rng_state(state)
        gum = -torch.empty_like(logits, memory_format=torch.legacy_contiguous_format).exponential_().log()
        out[f"{tag}_start"] = start.numpy(); out[f"{tag}_end"] = end.numpy()
        out[f"{tag}_gumbel"] = gum.numpy(); out[f"{tag}_index"] = index.numpy()
    np.savez_compressed(os.path.join(GOLD, "gumbel.npz"), **out)


def gen_transform():
    from src.gadgets import transforms as T
    u8 = syn.synthetic_frames_u8(3, 32, seed=5)
    u8[0, 0, 0] = torch.tensor([0, 255, 128], dtype=torch.uint8)
    ref = T.NormalizeVideo((0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711))(
        T.ToTensorVideo()(u8)).permute(1, 0, 2, 3).contiguous()
    # full chain from a float CTHW clip of another size (resize -> uint8 truncation)
    g = torch.Generator().manual_seed(9)
    clip = torch.rand(3, 2, 20, 28, generator=g) * 255.0
    chain = T.ToTHWC()(T.ToUint8()(T.ResizeVideo(16)(clip)))
    np.savez_compressed(os.path.join(GOLD, "transform.npz"), u8=u8.numpy(), norm=ref.numpy(), clip=clip.numpy(),
                        clip_u8=chain.contiguous().numpy())


def gen_resize():
    """N3: the reference transform chain on decoded (integer-valued) frames, as eval/utils/builder_utils.py:88-107,117-124
    feeds it: u8 [T,H,W,3] -> permute(3,0,1,2).float() -> ResizeVideo -> ToUint8 -> ToTHWC (-> ToTensorVideo -> NormalizeVideo)."""
    from src.gadgets import transforms as T
    rng = np.random.default_rng(21)
    out = {}
    for tag, (t, h, w, s) in dict(down=(2, 45, 80, 32), up=(2, 20, 28, 32), mixed=(1, 64, 24, 32), same=(1, 32, 32, 32)).items():
        u8 = rng.integers(0, 256, size=(t, h, w, 3), dtype=np.uint8)
        u8[0, : h // 2] = 200                                    # flat region: truncation is sensitive to the last fp32 bit
        clip = torch.from_numpy(u8).permute(3, 0, 1, 2).float()
        r = T.ToTHWC()(T.ToUint8()(T.ResizeVideo(s)(clip))).contiguous()
        full = T.NormalizeVideo((0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711))(
            T.ToTensorVideo()(r)).permute(1, 0, 2, 3).contiguous()
        out[f"{tag}_in"], out[f"{tag}_u8"], out[f"{tag}_norm"] = u8, r.numpy(), full.numpy()
    np.savez_compressed(os.path.join(GOLD, "resize.npz"), **out)


def gen_raft():
    """N1: the reference RAFT class (xraft.py) with the seeded synthetic state dict, eval mode, 20 iterations."""
    from src.models.components.xraft import RAFT
    from oracle import cases
    m = RAFT().eval()
    sd = syn.raft_state_dict()
    ref_sd = m.state_dict()
    assert set(ref_sd) == set(sd) and all(ref_sd[k].shape == sd[k].shape for k in sd)
    m.load_state_dict(sd)
    img1, img2 = cases.raft_pair()
    with torch.no_grad():
        flow = m(img1, img2, iters=20, test_mode=True)
        fmap1, _ = m.fnet([2 * (img1 / 255.0) - 1.0, 2 * (img2 / 255.0) - 1.0])
    np.savez_compressed(os.path.join(GOLD, "raft.npz"), flow=flow[:, :, ::4, ::4].numpy(), flow_absmax=float(flow.abs().max()),
                        flow_std=float(flow.std()), fmap1=fmap1[:, ::8, ::2, ::2].numpy())


def _summ(t: torch.Tensor, stride: int = 7):
    t = t.detach().float()
    return dict(shape=list(t.shape), mean=float(t.mean()), std=float(t.std()), absmax=float(t.abs().max())), \
        t.flatten()[::stride].clone().numpy()


def gen_models():
    reference_shim.install()
    from transformers import BertConfig, Blip2QFormerConfig, Blip2VisionConfig, InstructBlipQFormerConfig
    from src.models.components.xblip2 import Blip2QFormerModel, Blip2VisionModel
    from src.models.components.xinstructblip import InstructBlipQFormerModel
    from src.models.components.xropebert import RopeBertModel
    meta, arrays = {}, {}
    torch.manual_seed(0)

    def load(model, sd):
        missing, unexpected = model.load_state_dict(sd, strict=False)
        assert not unexpected, unexpected
        assert not [m for m in missing if "position_ids" not in m], missing
        return model.eval()

    # ---- ViT: full width, 2 layers (fast) and the full 39 layers on 1 frame
    for tag, layers, nf in (("vit2", 2, 2), ("vit39", 39, 1)):
        cfg = syn.VisionCfg(layers=layers)
        sd = syn.vision_state_dict(cfg)
        m = load(Blip2VisionModel(Blip2VisionConfig(num_hidden_layers=layers)), sd)
        assert abs(m.config.layer_norm_eps - cfg.eps) < 1e-12
        u8 = syn.synthetic_frames_u8(nf, seed=11)
        from oracle.transform_ref import normalize_u8_thwc
        px = torch.from_numpy(normalize_u8_thwc(u8.numpy()))
        with torch.no_grad():
            out = m(pixel_values=px, return_dict=True)
        meta[tag + "_last"], arrays[tag + "_last"] = _summ(out.last_hidden_state)
        meta[tag + "_pool"], arrays[tag + "_pool"] = _summ(out.pooler_output, 1)
        if tag == "vit2":
            img = out.last_hidden_state
    # ---- BLIP-2 Q-Former (12 layers) on the 2 frames above + projection
    qcfg = syn.QFormerCfg()
    qsd = syn.qformer_state_dict(qcfg)
    hsd = syn.head_state_dict()
    qm = load(Blip2QFormerModel(Blip2QFormerConfig()), qsd)
    assert abs(qm.config.layer_norm_eps - 1e-12) < 1e-20 and qm.config.cross_attention_frequency == 2
    qt = hsd["query_tokens"].expand(img.shape[0], -1, -1)
    with torch.no_grad():
        qo = qm(query_embeds=qt, encoder_hidden_states=img,
                encoder_attention_mask=torch.ones(img.shape[:2], dtype=torch.long))[0]
        proj = torch.nn.functional.linear(qo, hsd["language_projection.weight"], hsd["language_projection.bias"])
    meta["qf_out"], arrays["qf_out"] = _summ(qo, 3)
    meta["qf_proj"], arrays["qf_proj"] = _summ(proj.reshape(1, -1, proj.shape[-1]), 5)
    # ---- InstructBLIP Q-Former with padded instruction text
    icfg = syn.QFormerCfg(instruct=True)
    isd = syn.qformer_state_dict(icfg)
    im = load(InstructBlipQFormerModel(InstructBlipQFormerConfig()), isd)
    g = torch.Generator().manual_seed(21)
    ids = torch.randint(1000, 30522, (1, 12), generator=g).repeat_interleave(img.shape[0], 0)
    tmask = torch.ones(img.shape[0], 12, dtype=torch.long); tmask[:, 9:] = 0
    amask = torch.cat([torch.ones(img.shape[0], 32, dtype=torch.long), tmask], dim=1)
    with torch.no_grad():
        io = im(input_ids=ids, attention_mask=amask, query_embeds=qt, encoder_hidden_states=img,
                encoder_attention_mask=torch.ones(img.shape[:2], dtype=torch.long), return_dict=True).last_hidden_state
    arrays["iq_ids"] = ids.numpy(); arrays["iq_mask"] = amask.numpy()
    meta["iq_out"], arrays["iq_out"] = _summ(io, 3)
    # ---- sampler, both modes, padded flow + padded question
    scfg = syn.SamplerCfg()
    ssd = syn.sampler_state_dict(scfg)
    sm = load(RopeBertModel(BertConfig(fusion_layer=6, encoder_width=768)), ssd)
    b, lof, tq = 2, 12, 9
    of = syn.synthetic_flow(b, lof, seed=31)
    of_len = [12, 7]
    of_mask = torch.zeros(b, lof + 2, dtype=torch.long)
    for j, n in enumerate(of_len):
        of_mask[j, : n + 2] = 1
    qids = torch.randint(1000, 30522, (b, tq), generator=g)
    qmask = torch.ones(b, tq, dtype=torch.long); qmask[1, 6:] = 0
    arrays["s_qids"] = qids.numpy(); arrays["s_qmask"] = qmask.numpy(); arrays["s_ofmask"] = of_mask.numpy()
    for mode in ("fusion", "multi_modal"):
        with torch.no_grad():
            seq, logits = sm(encoder_embeds=of, attention_mask=of_mask, encoder_hidden_states=qids,
                             encoder_attention_mask=qmask, mode=mode)
        meta[f"s_{mode}_seq"], arrays[f"s_{mode}_seq"] = _summ(seq, 3)
        meta[f"s_{mode}_logits"], arrays[f"s_{mode}_logits"] = _summ(logits, 1)
    with torch.no_grad():
        emb = sm.temporal_embeddings(of, of_mask)
    meta["s_ofembed"], arrays["s_ofembed"] = _summ(emb, 1)
    np.savez_compressed(os.path.join(GOLD, "models.npz"), **arrays)
    with open(os.path.join(GOLD, "models_meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print(json.dumps({k: v for k, v in meta.items()}, indent=0)[:2000])


def gen_task_module():
    """(b) the boundary at its own seam: the UNMODIFIED `LSTPSFModule.eval_forward` methods of the reference task modules
    (src/models/LSTP_SF_blip2_module.py:382-506 BLIP-2, src/models/LSTP_SF_module.py:456-580 InstructBLIP) run on the
    reference's own container / sampler classes built from real HF config objects, CPU fp32.  The stand-in language
    model records what `generate` receives; that (inputs_embeds, attention_mask) and the returned ids are the fixture
    `tests/test_boundary_gpu.py` holds the drop-in classes to."""
    import importlib
    reference_shim.install()
    sys.path.insert(1, os.path.join(ROOT, "tests"))
    import reference_stubs
    reference_stubs.install()
    from tests import task_module_case as tc
    arrays, meta = {}, {}
    for tag, modname, instruct in (("blip2", "src.models.LSTP_SF_blip2_module", False),
                                   ("instruct", "src.models.LSTP_SF_module", True)):
        mod = importlib.import_module(modname)
        assert mod.__file__.startswith(REF), mod.__file__
        cfg, bert_cfg = tc.hf_configs(instruct)
        container_cls = mod.InstructBlipForConditionalGeneration if instruct else mod.Blip2ForConditionalGeneration
        assert container_cls.__module__.startswith("src.models.components") and \
            sys.modules[container_cls.__module__].__file__.startswith(REF)
        sds = tc.state_dicts(instruct)
        model = container_cls(cfg)
        res = model.load_state_dict(tc.container_state_dict(sds), strict=False)
        assert not res.unexpected_keys, res.unexpected_keys
        assert all(k.startswith(("language_model.", "temporal_projection.")) or "position_ids" in k for k in res.missing_keys), \
            res.missing_keys
        lm = tc.CaptureLM()
        model.language_model = lm
        sampler = mod.RopeBertModel(config=bert_cfg)
        res = sampler.load_state_dict(sds["sampler"], strict=False)
        assert not res.unexpected_keys and all("position_ids" in k for k in res.missing_keys), res
        task = mod.LSTPSFModule.__new__(mod.LSTPSFModule)            # __init__ needs hub checkpoints; attributes suffice
        torch.nn.Module.__init__(task)
        task.model, task.temporal_encoder = model.eval(), sampler.eval()
        task.generate_configs = dict(max_new_tokens=3)
        batch, gumbel = tc.make_batch(instruct)
        with tc.fixed_gumbel(gumbel):
            outputs = mod.LSTPSFModule.eval_forward(task, batch)
        emb, am = lm.seen["inputs_embeds"], lm.seen["attention_mask"]
        assert lm.seen["kwargs"] == dict(max_new_tokens=3)
        arrays[tag + "_inputs_embeds"] = emb.numpy().astype(np.float32)
        arrays[tag + "_attention_mask"] = am.numpy()
        arrays[tag + "_outputs"] = outputs.numpy()
        meta[tag] = dict(shape=list(emb.shape), absmax=float(emb.abs().max()), std=float(emb.std()))
        print(tag, meta[tag], outputs.tolist())
    np.savez_compressed(os.path.join(GOLD, "task_module.npz"), **arrays)
    with open(os.path.join(GOLD, "task_module_meta.json"), "w") as f:
        json.dump(meta, f, indent=1)


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    reference_shim.install()
    which = sys.argv[1:] or ["select", "pseudo_span", "gumbel", "transform", "resize", "raft", "models", "task_module"]
    for w in which:
        globals()["gen_" + w]()
        print("generated", w)
