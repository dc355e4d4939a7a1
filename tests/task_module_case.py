"""The seeded end-to-end case shared by `oracle/gen_golden.py task_module` (which runs the UNMODIFIED reference
`eval_forward` methods on the reference classes, CPU fp32, and stores what they hand to `language_model.generate`) and
`tests/test_boundary_gpu.py` (which drives the drop-in classes imported through `src.models.components.*` with the same
HF config objects, state dicts, batch and Gumbel noise).  Test infrastructure only."""
from __future__ import annotations

import torch
from torch import nn

from videotgb_b200 import synthetic as syn

LLM_HIDDEN = 256
VOCAB = 512
VIT_LAYERS = 2
B, N, LOF, TQ, K = 2, 8, 10, 7, 4
OF_LENGTHS = [10, 6]


def hf_configs(instruct: bool):
    """Real HF config objects, as the reference builds its models from (Blip2Config / InstructBlipConfig, BertConfig)."""
    import transformers as tf
    vis = dict(num_hidden_layers=VIT_LAYERS)
    if instruct:
        text = tf.LlamaConfig(hidden_size=LLM_HIDDEN, intermediate_size=512, num_hidden_layers=1, num_attention_heads=4,
                              num_key_value_heads=4, vocab_size=VOCAB, architectures=["LLaMAForCausalLM"])
        cfg = tf.InstructBlipConfig(vision_config=vis, qformer_config={}, text_config=text.to_dict(), num_query_tokens=32)
    else:
        text = tf.T5Config(d_model=LLM_HIDDEN, d_ff=512, d_kv=64, num_layers=1, num_decoder_layers=1, num_heads=4,
                           vocab_size=VOCAB, architectures=["T5ForConditionalGeneration"])
        cfg = tf.Blip2Config(vision_config=vis, qformer_config={}, text_config=text.to_dict(), num_query_tokens=32)
    return cfg, tf.BertConfig(fusion_layer=6, encoder_width=768)


class CaptureLM(nn.Module):
    """Stands in for the unchanged HF language model: records what `generate` receives."""

    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(77)
        self.embed = nn.Embedding(VOCAB, LLM_HIDDEN)
        with torch.no_grad():
            self.embed.weight.copy_(torch.randn(VOCAB, LLM_HIDDEN, generator=g) * 0.02)
        self.seen = None

    def get_input_embeddings(self):
        return self.embed

    def generate(self, inputs_embeds=None, attention_mask=None, **kw):
        self.seen = dict(inputs_embeds=inputs_embeds, attention_mask=attention_mask, kwargs=kw)
        b = inputs_embeds.shape[0]
        return torch.tensor([[5, 0, 9, 0]] * b, dtype=torch.long, device=inputs_embeds.device)


def state_dicts(instruct: bool):
    qcfg = syn.QFormerCfg(instruct=instruct)
    return dict(vision=syn.vision_state_dict(syn.VisionCfg(layers=VIT_LAYERS)), qformer=syn.qformer_state_dict(qcfg),
                sampler=syn.sampler_state_dict(), head=syn.head_state_dict(llm_hidden=LLM_HIDDEN))


def container_state_dict(sds) -> dict:
    """Keys as in a reference checkpoint of the container (`vision_model.*`, `qformer.*`, `query_tokens`, ...)."""
    out = {"vision_model." + k: v for k, v in sds["vision"].items()}
    out.update({"qformer." + k: v for k, v in sds["qformer"].items()})
    out.update(sds["head"])
    return out


def make_batch(instruct: bool, device="cpu"):
    """Batch dict with the keys the reference collate produces (videoinstruct_dataset.py:147-192)."""
    from oracle.transform_ref import normalize_u8_thwc
    g = torch.Generator().manual_seed(4242)
    u8 = torch.randint(0, 256, (B * N, 224, 224, 3), generator=g, dtype=torch.uint8)
    frames = torch.from_numpy(normalize_u8_thwc(u8.numpy()))                  # [B*N, 3, 224, 224] fp32 (A0)
    of = syn.synthetic_flow(B, LOF, seed=43)
    of_mask = torch.zeros(B, LOF + 2, dtype=torch.long)
    for j, n in enumerate(OF_LENGTHS):
        of_mask[j, : n + 2] = 1
    batch = dict(frames=frames, of=of, of_mask=of_mask, of_lengths=list(OF_LENGTHS), nframe=K,
                 sampler_question=torch.randint(1000, 30522, (B, TQ), generator=g),
                 sampler_question_attention_mask=torch.ones(B, TQ, dtype=torch.long),
                 question=torch.randint(1, VOCAB, (B, 5), generator=g),
                 question_attention_mask=torch.ones(B, 5, dtype=torch.long),
                 answer=torch.zeros(B, 1, dtype=torch.long))
    batch["sampler_question_attention_mask"][1, 5:] = 0
    batch["question_attention_mask"][1, 3:] = 0
    if instruct:
        batch["qformer_text"] = torch.randint(1000, 30522, (B, 9), generator=g)
        m = torch.ones(B, 9, dtype=torch.long)
        m[0, 6:] = 0
        batch["qformer_text_attention_mask"] = m
    gumbel = [-torch.empty(2 * B, LOF, 1).exponential_(generator=g).log() for _ in range(2)]
    if device != "cpu":
        batch = {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in batch.items()}
        gumbel = [x.to(device) for x in gumbel]
    return batch, gumbel


class fixed_gumbel:
    """Context manager: `F.gumbel_softmax` draws its noise from the given list (one tensor per call) instead of the
    RNG -- the arithmetic around the noise is torch's own (`softmax((logits + g) / tau, dim)`, hard=False)."""

    def __init__(self, noise):
        self.noise = list(noise)

    def __enter__(self):
        import torch.nn.functional as F
        self._orig = F.gumbel_softmax
        it = iter(self.noise)

        def gumbel_softmax(logits, tau=1, hard=False, eps=1e-10, dim=-1):
            assert not hard
            g = next(it).to(logits.device).reshape(logits.shape)
            return ((logits + g) / tau).softmax(dim)

        F.gumbel_softmax = gumbel_softmax
        return self

    def __exit__(self, *exc):
        import torch.nn.functional as F
        F.gumbel_softmax = self._orig
        return False
