"""GPU: the drop-in boundary at its own seam.

The drop-in classes are imported THROUGH the reference's import paths (`src.models.components.*`), built from real HF
config objects (`Blip2Config` / `InstructBlipConfig`, `BertConfig(fusion_layer=6, encoder_width=768)`), filled with the
same seeded state dicts and driven with the same batch and Gumbel noise as the UNMODIFIED reference `eval_forward`
methods were when `oracle/gen_golden.py task_module` recorded what they pass to `language_model.generate`
(tests/golden/task_module.npz).  Tolerance: north_star 2e-2 / cosine 0.999 on the visual tokens; the question-embedding
rows, the attention mask and the post-processed ids are exact.

Where a reference checkout is available next to a GPU (TGB_REFERENCE_ROOT, default /root/reference -- never on the
driver's GPU box), the reference's own `eval_forward` source additionally runs, unmodified, AGAINST the drop-in objects."""
import json
import os
import sys

import numpy as np
import pytest
import torch

from tests import task_module_case as tc
from tests.parity_utils import assert_parity

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
REFERENCE = os.environ.get("TGB_REFERENCE_ROOT", "/root/reference")
DEV = "cuda"


def _build(instruct):
    from src.models.components.xblip2 import Blip2ForConditionalGeneration
    from src.models.components.xinstructblip import InstructBlipForConditionalGeneration
    from src.models.components.xropebert import RopeBertModel
    cfg, bert_cfg = tc.hf_configs(instruct)
    lm = tc.CaptureLM()
    cls = InstructBlipForConditionalGeneration if instruct else Blip2ForConditionalGeneration
    model = cls(cfg, language_model=lm)
    sds = tc.state_dicts(instruct)
    res = model.load_state_dict(tc.container_state_dict(sds), strict=False)
    assert not res.unexpected_keys
    assert all(k.startswith(("language_model.", "temporal_projection.")) for k in res.missing_keys), res.missing_keys
    sampler = RopeBertModel(config=bert_cfg)
    sampler.load_state_dict(sds["sampler"])
    return model.to(DEV).eval(), sampler.to(DEV).eval(), lm


def _check(tag, lm, outputs):
    gold = np.load(os.path.join(GOLDEN, "task_module.npz"))
    with open(os.path.join(GOLDEN, "task_module_meta.json")) as f:
        meta = json.load(f)[tag]
    emb = lm.seen["inputs_embeds"].float().cpu()
    want = torch.from_numpy(gold[tag + "_inputs_embeds"])
    assert list(emb.shape) == meta["shape"]
    nv = tc.K * 32
    r, c = assert_parity(emb[:, :nv], want[:, :nv], f"{tag}: visual tokens vs the reference eval_forward")
    assert torch.equal(emb[:, nv:], want[:, nv:]), "question embedding rows must be exact copies"
    assert torch.equal(lm.seen["attention_mask"].cpu(), torch.from_numpy(gold[tag + "_attention_mask"]))
    assert torch.equal(outputs.cpu(), torch.from_numpy(gold[tag + "_outputs"]))
    assert lm.seen["kwargs"] == dict(max_new_tokens=3)
    print(f"{tag}: relerr {r:.3e} cosine {c:.6f}")


@pytest.mark.parametrize("instruct", [False, True], ids=["blip2", "instructblip"])
def test_eval_forward_through_reference_import_paths(instruct):
    from videotgb_b200.pipeline import VideoTGBEncoder
    model, sampler, lm = _build(instruct)
    enc = VideoTGBEncoder.from_modules(model, sampler)
    batch, gumbel = tc.make_batch(instruct, DEV)
    outputs = enc.eval_forward(batch, generate_configs=dict(max_new_tokens=3), mode="fusion", gumbel=gumbel)
    _check("instruct" if instruct else "blip2", lm, outputs)


def test_module_protocol_with_hf_configs():
    """The call sites of SURVEY.md section 8b one by one, on objects built from HF configs."""
    model, sampler, _ = _build(False)
    batch, _ = tc.make_batch(False, DEV)
    of_feat, of_logits = sampler(encoder_embeds=batch["of"], attention_mask=batch["of_mask"],
                                 encoder_hidden_states=batch["sampler_question"],
                                 encoder_attention_mask=batch["sampler_question_attention_mask"], mode="fusion")
    assert of_feat.shape == (tc.B, tc.LOF + 2, 768) and of_logits.shape == (tc.B, tc.LOF, 2)
    px = batch["frames"][:3]
    out = model.vision_model(pixel_values=px, return_dict=True)
    assert out.last_hidden_state.shape == (3, 257, 1408) and out.pooler_output.shape == (3, 1408)
    assert torch.equal(model.vision_model(pixel_values=px)[0], out.last_hidden_state)       # config.return_dict picked up
    q = model.query_tokens.expand(3, -1, -1)
    qo = model.qformer(query_embeds=q, encoder_hidden_states=out.last_hidden_state,
                       encoder_attention_mask=torch.ones(3, 257, dtype=torch.long, device=DEV))[0]
    assert qo.shape == (3, 32, 768)
    y = model.language_projection(qo)
    assert y.shape == (3, 32, tc.LLM_HIDDEN)
    assert model.get_input_embeddings()(batch["question"]).shape == (tc.B, 5, tc.LLM_HIDDEN)
    assert model.config.text_config.architectures[0] == "T5ForConditionalGeneration"
    assert model.config.use_decoder_only_language_model is False
    frozen = [p.requires_grad_(False) for p in model.vision_model.parameters()]              # freeze_weights(), :664-666
    assert len(frozen) > 0


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "src", "models")),
                    reason="no reference checkout on this box (the fixture test above covers the same path)")
@pytest.mark.parametrize("instruct", [False, True], ids=["blip2", "instructblip"])
def test_unmodified_reference_eval_forward_on_drop_ins(instruct):
    import importlib
    if REFERENCE not in sys.path:
        sys.path.append(REFERENCE)                     # AFTER this repo: the overlay `src` package wins
    from tests import reference_stubs
    reference_stubs.install()
    mod = importlib.import_module("src.models.LSTP_SF_module" if instruct else "src.models.LSTP_SF_blip2_module")
    assert mod.__file__.startswith(REFERENCE) and mod.RopeBertModel.__module__ == "videotgb_b200.sampler"
    model, sampler, lm = _build(instruct)
    task = mod.LSTPSFModule.__new__(mod.LSTPSFModule)
    torch.nn.Module.__init__(task)
    task.model, task.temporal_encoder, task.generate_configs = model, sampler, dict(max_new_tokens=3)
    batch, gumbel = tc.make_batch(instruct, DEV)
    with tc.fixed_gumbel(gumbel):
        outputs = mod.LSTPSFModule.eval_forward(task, batch)
    _check("instruct" if instruct else "blip2", lm, outputs)
