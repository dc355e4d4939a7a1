"""CPU: the drop-in boundary at its own seam -- what can be checked without a device.

* the `src` overlay package: with this repo in front of the reference checkout on PYTHONPATH the reference's UNMODIFIED
  task modules import, and the names they import from `.components.*` are the B200 classes;
* `from_pretrained` of the containers and of the sampler from HF-layout checkpoint directories written on the fly;
* constructor pickup from real HF config objects."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.environ.get("TGB_REFERENCE_ROOT", "/root/reference")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "src", "models")), reason="reference checkout not present")
def test_reference_task_modules_import_the_drop_ins():
    code = r"""
import sys, json
sys.path.insert(0, sys.argv[1] + "/tests")
import reference_stubs; reference_stubs.install()
import src.models.LSTP_SF_blip2_module as a, src.models.LSTP_SF_module as b, src.models.LSTP_module as c
out = {}
for tag, m in (("blip2", a), ("instruct", b), ("lstp", c)):
    out[tag] = dict(file=m.__file__, **{n: getattr(m, n).__module__ for n in
                    ("RopeBertModel", "RAFT", "InputPadder", "dp_state_to_normal", "TemporalOFEmbedding", "rouge_n")
                    if hasattr(m, n)})
out["blip2"]["container"] = a.Blip2ForConditionalGeneration.__module__
out["instruct"]["container"] = b.InstructBlipForConditionalGeneration.__module__
out["has_from_pretrained"] = all(hasattr(x, "from_pretrained") for x in
                                 (a.Blip2ForConditionalGeneration, b.InstructBlipForConditionalGeneration, a.RopeBertModel))
print(json.dumps(out))
"""
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + REFERENCE)
    r = subprocess.run([sys.executable, "-c", code, ROOT], capture_output=True, text=True, env=env, cwd="/tmp", timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    for tag in ("blip2", "instruct", "lstp"):
        o = out[tag]
        assert o["file"].startswith(REFERENCE), o                     # the task module itself is the reference's file
        assert o["RopeBertModel"] == "videotgb_b200.sampler"
        assert o["RAFT"] == "videotgb_b200.raft" and o["InputPadder"] == "videotgb_b200.raft"
        assert o["dp_state_to_normal"] == "videotgb_b200.raft"
        assert o["TemporalOFEmbedding"] == "src.models.components.xsampler"      # not replaced: falls through
        assert o.get("rouge_n", "src.gadgets.my_metrics") == "src.gadgets.my_metrics"
    assert out["blip2"]["container"] == "videotgb_b200.containers"
    assert out["instruct"]["container"] == "videotgb_b200.containers"
    assert out["has_from_pretrained"]


def test_shim_modules_export_what_the_task_modules_import():
    """Without the reference on the path: every name of LSTP_SF_blip2_module.py:20-25 / eval/utils/model.py:12-17
    that this repo replaces resolves."""
    from src.models.components.xblip2 import Blip2ForConditionalGeneration, Blip2QFormerModel, Blip2VisionModel  # noqa: F401
    from src.models.components.xinstructblip import (InstructBlipForConditionalGeneration, InstructBlipQFormerModel,  # noqa: F401
                                                     InstructBlipVisionModel)
    from src.models.components.xropebert import RopeBertModel
    from src.models.components.xraft import RAFT, InputPadder  # noqa: F401
    from src.models.components.raft_utils.utils import dp_state_to_normal
    assert callable(RopeBertModel.from_pretrained) and callable(Blip2ForConditionalGeneration.from_pretrained)
    sd = {"module.fnet.conv1.weight": 1, "module.update_block.module.x": 2, "fnet.orphan": 3}
    assert dp_state_to_normal(sd) == {"fnet.conv1.weight": 1, "update_block.x": 2}      # raft_utils/utils.py:85-91
    for dims, mode, want in (((3, 224, 224), "sintel", [0, 0, 0, 0]), ((3, 436, 1024), "sintel", [0, 0, 2, 2]),
                             ((3, 375, 1242), "kitti", [3, 3, 0, 1]), ((1, 3, 90, 161), "sintel", [3, 4, 3, 3])):
        assert InputPadder(dims, mode)._pad == want, (dims, mode)


def _tiny_configs(instruct):
    from tests.task_module_case import hf_configs
    return hf_configs(instruct)


@pytest.mark.parametrize("instruct", [False, True])
def test_container_from_hf_config_and_from_pretrained(tmp_path, instruct):
    from safetensors.torch import save_file
    from src.models.components import xblip2, xinstructblip
    from tests import task_module_case as tc
    cls = xinstructblip.InstructBlipForConditionalGeneration if instruct else xblip2.Blip2ForConditionalGeneration
    cfg, _ = _tiny_configs(instruct)
    # constructor pickup from the real HF config objects
    m = cls(cfg, language_model=False)
    assert m.vision_model.layers_n == tc.VIT_LAYERS and m.vision_model.eps == cfg.vision_config.layer_norm_eps
    assert m.vision_model.hidden == cfg.vision_config.hidden_size and m.vision_model.mlp == cfg.vision_config.intermediate_size
    assert m.qformer.eps == cfg.qformer_config.layer_norm_eps and m.qformer.cross_freq == cfg.qformer_config.cross_attention_frequency
    assert m.qformer.enc_hidden == cfg.qformer_config.encoder_hidden_size
    assert tuple(m.query_tokens.shape) == (1, cfg.num_query_tokens, cfg.qformer_config.hidden_size)
    assert m.language_projection.weight.shape == (tc.LLM_HIDDEN, 768) and m.temporal_projection.weight.shape == (tc.LLM_HIDDEN, 768)
    with pytest.raises(RuntimeError):
        m.get_input_embeddings()
    # the reference builds the unchanged HF language model from text_config; so does the container by default
    full = cls(cfg)
    assert type(full.language_model).__name__ == ("LlamaForCausalLM" if instruct else "T5ForConditionalGeneration")
    assert full.get_input_embeddings() is full.language_model.get_input_embeddings()
    # HF-layout checkpoint directory -> from_pretrained
    sds = tc.state_dicts(instruct)
    sd = tc.container_state_dict(sds)
    sd.update({"language_model." + k: v.clone() for k, v in full.language_model.state_dict().items()})
    cfg.save_pretrained(tmp_path)
    half = len(sd) // 2
    keys = sorted(sd)
    shards = {"model-00001-of-00002.safetensors": keys[:half], "model-00002-of-00002.safetensors": keys[half:]}
    for fn, ks in shards.items():
        save_file({k: sd[k].contiguous() for k in ks}, os.path.join(tmp_path, fn))
    with open(os.path.join(tmp_path, "model.safetensors.index.json"), "w") as f:
        json.dump({"weight_map": {k: fn for fn, ks in shards.items() for k in ks}}, f)
    got = cls.from_pretrained(str(tmp_path), language_model=False)
    own = got.state_dict()
    for k, v in tc.container_state_dict(sds).items():
        assert torch.equal(own[k], v), k
    assert not any(k.startswith("language_model.") for k in own)
    got_full = cls.from_pretrained(str(tmp_path))
    for k, v in full.language_model.state_dict().items():
        assert torch.equal(got_full.language_model.state_dict()[k], v), k
    assert not got.training


def test_sampler_from_pretrained_bert_checkpoint(tmp_path):
    """`RopeBertModel.from_pretrained("bert-base-uncased", config=BertConfig(fusion_layer=6, encoder_width=768))`
    (LSTP_SF_blip2_module.py:88-89): the `bert.` prefix is stripped, legacy gamma/beta renamed, sampler-only parameters
    keep their initialisation."""
    from transformers import BertConfig
    from src.models.components.xropebert import RopeBertModel
    cfg = BertConfig(fusion_layer=1, encoder_width=64, hidden_size=64, num_hidden_layers=2, num_attention_heads=4,
                     intermediate_size=128, vocab_size=100, max_position_embeddings=40)
    g = torch.Generator().manual_seed(3)
    ref = RopeBertModel(cfg)
    bert_sd = {}
    for k, v in ref.state_dict().items():
        if k.startswith(("temporal_embeddings.", "mrc_head.", "encoder.embed_positions", "encoder.c_embed_positions")) \
                or "crossattention" in k:
            continue
        t = torch.randn(v.shape, generator=g) if v.is_floating_point() else v.clone()
        bert_sd["bert." + k.replace("LayerNorm.weight", "LayerNorm.gamma").replace("LayerNorm.bias", "LayerNorm.beta")] = t
    bert_sd["cls.predictions.bias"] = torch.zeros(100)
    bert_sd["bert.pooler.dense.weight"] = torch.zeros(64, 64)
    torch.save(bert_sd, os.path.join(tmp_path, "pytorch_model.bin"))
    m = RopeBertModel.from_pretrained(str(tmp_path), config=cfg)
    own = m.state_dict()
    assert torch.equal(own["encoder.layer.1.attention.self.query.weight"], bert_sd["bert.encoder.layer.1.attention.self.query.weight"])
    assert torch.equal(own["embeddings.LayerNorm.weight"], bert_sd["bert.embeddings.LayerNorm.gamma"])
    assert torch.equal(own["encoder.layer.0.output.LayerNorm.bias"], bert_sd["bert.encoder.layer.0.output.LayerNorm.beta"])
    assert any(k.startswith("temporal_embeddings.") for k in m.missing_keys) and "mrc_head.weight" in m.missing_keys
    assert any("crossattention" in k for k in m.missing_keys)
    assert not m.training and m.fusion_layer == 1 and m.enc_width == 64
    # a VideoTGB checkpoint (Lightning `state_dict` with the `temporal_encoder.` prefix) loads through the same call
    torch.save({"state_dict": {"temporal_encoder." + k: v for k, v in ref.state_dict().items()}}, os.path.join(tmp_path, "tgb.ckpt"))
    m2 = RopeBertModel.from_pretrained(os.path.join(tmp_path, "tgb.ckpt"), config=cfg)
    for k, v in ref.state_dict().items():
        assert torch.equal(m2.state_dict()[k], v), k
