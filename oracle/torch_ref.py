"""TEST INFRASTRUCTURE ONLY -- plain-torch restatement of the floating-point stages of the hot path.

Functional (state-dict in, tensor out) restatement of the reference modules, one function per
module, each citing the reference lines it follows.  Written with ordinary torch ops so that

* in fp32 it is the oracle the golden fixtures pin (``tests/test_oracle_golden.py``), and
* wrapped in ``torch.autocast("cuda", torch.bfloat16)`` on the GPU box it reproduces the
  reference's own mixed-precision path (Lightning ``precision: bf16``: matmul/conv in bf16,
  softmax/LayerNorm in fp32, fp32 master weights), which is the 2e-2 parity target.

Dropout is identity (parity is defined in eval mode).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def _lin(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def _ln(sd, name, x, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], eps)


def _gelu(x):
    return F.gelu(x)                     # HF ACT2FN["gelu"]: exact erf GELU


# --------------------------------------------------------------------------------------------
# A10-A13  EVA-ViT-g  (xblip2.py:89-117 embeddings, :120-198 attention, :202-214 MLP,
#                      :218-263 layer, :500-563 model)
# --------------------------------------------------------------------------------------------
def vit_embeddings(sd, pixel_values, patch=14):
    w = sd["embeddings.patch_embedding.weight"]
    x = F.conv2d(pixel_values.to(w.dtype), w, sd["embeddings.patch_embedding.bias"], stride=patch)
    x = x.flatten(2).transpose(1, 2)                                   # [F, 256, D]
    cls = sd["embeddings.class_embedding"].expand(x.shape[0], 1, -1).to(w.dtype)
    x = torch.cat([cls, x], dim=1)
    return x + sd["embeddings.position_embedding"][:, : x.shape[1], :].to(w.dtype)


def vit_attention(sd, p, x, heads):
    b, s, d = x.shape
    qkv = _lin(sd, p + "self_attn.qkv", x)
    qkv = qkv.reshape(b, s, 3, heads, d // heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    scores = torch.matmul(q, k.transpose(-1, -2)) * ((d // heads) ** -0.5)
    probs = F.softmax(scores, dim=-1)
    ctx = torch.matmul(probs, v).permute(0, 2, 1, 3).reshape(b, s, d)
    return _lin(sd, p + "self_attn.projection", ctx)


def vit_layer(sd, i, x, heads, eps):
    p = f"encoder.layers.{i}."
    x = vit_attention(sd, p, _ln(sd, p + "layer_norm1", x, eps), heads) + x
    h = _ln(sd, p + "layer_norm2", x, eps)
    h = _lin(sd, p + "mlp.fc2", _gelu(_lin(sd, p + "mlp.fc1", h)))
    return h + x


def vit_forward(sd, pixel_values, layers=39, heads=16, eps=1e-6, patch=14):
    """-> (last_hidden_state [F,257,D], pooler_output [F,D])  (xblip2.py:517-560)."""
    x = vit_embeddings(sd, pixel_values, patch)
    for i in range(layers):
        x = vit_layer(sd, i, x, heads, eps)
    last = _ln(sd, "post_layernorm", x, eps)
    pooled = _ln(sd, "post_layernorm", last[:, 0, :], eps)             # double LN, as the reference
    return last, pooled


# --------------------------------------------------------------------------------------------
# A14/A15  Q-Former  (xblip2.py:566-695 attention, :699-790 dense blocks, :793-890 layer,
#                     :1063-1174 model; InstructBLIP twin xinstructblip.py:565-694, :814-883,
#                     :1018-1046 embeddings, :1122-1242 model)
# --------------------------------------------------------------------------------------------
def _mha(sd, p, x_q, x_kv, add_mask, heads):
    b, sq, d = x_q.shape
    hd = d // heads

    def split(t):
        return t.view(b, -1, heads, hd).permute(0, 2, 1, 3)

    k = split(_lin(sd, p + "key", x_kv))
    v = split(_lin(sd, p + "value", x_kv))
    q = split(_lin(sd, p + "query", x_q))
    scores = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(hd)
    if add_mask is not None:
        scores = scores + add_mask
    probs = torch.softmax(scores, dim=-1)
    ctx = torch.matmul(probs, v).permute(0, 2, 1, 3).contiguous().view(b, sq, d)
    return ctx


def _bert_out(sd, p, h, resid, eps):
    return _ln(sd, p + "LayerNorm", _lin(sd, p + "dense", h) + resid, eps)


def qformer_forward(sd, query_embeds, encoder_hidden_states, encoder_attention_mask=None,
                    input_ids=None, attention_mask=None, layers=12, heads=12, cross_freq=2, eps=1e-12,
                    instruct=False):
    """-> last_hidden_state [F, 32(+Tt), 768].

    BLIP-2 (``instruct=False``): LayerNorm(query_embeds) then 12 post-LN layers.
    InstructBLIP: word+position embeddings of ``input_ids`` concatenated after the queries, shared
    self-attention with the ``(1-mask)*-10000`` padding mask, text rows use their own FFN and never
    cross-attend.
    """
    nq = query_embeds.shape[1]
    if instruct:
        x = query_embeds
        if input_ids is not None:
            t = sd["embeddings.word_embeddings.weight"][input_ids]
            t = t + sd["embeddings.position_embeddings.weight"][: input_ids.shape[1]].unsqueeze(0)
            x = torch.cat([query_embeds, t], dim=1)
        x = _ln(sd, "embeddings.layernorm", x, eps)
    else:
        x = _ln(sd, "layernorm", query_embeds, eps)
    b, s, _ = x.shape
    if attention_mask is None:
        attention_mask = torch.ones(b, s, device=x.device)
    self_mask = (1.0 - attention_mask[:, None, None, :].to(x.dtype)) * -10000.0
    if encoder_attention_mask is None:
        encoder_attention_mask = torch.ones(encoder_hidden_states.shape[:2], device=x.device)
    em = encoder_attention_mask[:, None, None, :].to(x.dtype)
    cross_mask = (1.0 - em) * torch.finfo(x.dtype).min               # HF invert_attention_mask
    for i in range(layers):
        p = f"encoder.layer.{i}."
        a = _mha(sd, p + "attention.attention.", x, x, self_mask, heads)
        a = _bert_out(sd, p + "attention.output.", a, x, eps)
        q_rows = a[:, :nq, :]
        if i % cross_freq == 0:
            c = _mha(sd, p + "crossattention.attention.", q_rows, encoder_hidden_states, cross_mask, heads)
            q_rows = _bert_out(sd, p + "crossattention.output.", c, q_rows, eps)
        out = _bert_out(sd, p + "output_query.", _gelu(_lin(sd, p + "intermediate_query.dense", q_rows)), q_rows, eps)
        if s > nq:
            t_rows = a[:, nq:, :]
            t_out = _bert_out(sd, p + "output.", _gelu(_lin(sd, p + "intermediate.dense", t_rows)), t_rows, eps)
            out = torch.cat([out, t_out], dim=1)
        x = out
    return x


# --------------------------------------------------------------------------------------------
# A16  language_projection + token assembly  (xblip2.py:1552; LSTP_SF_blip2_module.py:480-481;
#      mean-pool eval/utils/model.py:186-195)
# --------------------------------------------------------------------------------------------
def project_tokens(sd, query_output, batch, mode="concat", widths=None):
    if mode == "concat":
        y = _lin(sd, "language_projection", query_output)
        return y.reshape(batch, -1, y.shape[-1])
    if widths is None:
        widths = [query_output.shape[0] // batch] * batch
    pooled = torch.zeros((batch, query_output.shape[1], query_output.shape[2]), dtype=query_output.dtype,
                         device=query_output.device)
    idx = 0
    for i, w in enumerate(widths):
        pooled[i] = query_output[idx: idx + w].mean(0)
        idx += w
    return _lin(sd, "language_projection", pooled)


# --------------------------------------------------------------------------------------------
# A1-A6  temporal sampler (RoPE-BERT)  (xropebert.py:103-129 OF embedding, :190-208 text
#        embedding, :243-377 attention + RoPE, :450-538 layer, :596-711 encoder, :1048-1178 model)
# --------------------------------------------------------------------------------------------
def of_embedding(sd, of, of_mask, patch=16):
    b, l, c, h, w = of.shape
    p = "temporal_embeddings."
    x = F.conv2d(of.view(-1, c, h, w), sd[p + "projection.weight"], sd[p + "projection.bias"], stride=patch)
    x = x.flatten(2)                                                   # [B*L, 768, 196]
    x = F.linear(x, sd[p + "fc.weight"], sd[p + "fc.bias"]).view(b, l, -1)
    d = x.shape[-1]
    x = torch.cat([sd[p + "bos"].expand(b, 1, -1).to(x.dtype), x, torch.zeros(b, 1, d, device=x.device, dtype=x.dtype)], dim=1)
    ends = of_mask.sum(dim=1) - 1
    x[torch.arange(b), ends] = sd[p + "eos"].to(x.dtype)
    x = x + sd[p + "frame_pos_embed.weight"][: l + 2].unsqueeze(0)
    return F.layer_norm(x, (d,), sd[p + "ln.weight"], sd[p + "ln.bias"], 1e-5)     # torch default eps (:91)


def text_embedding(sd, ids, eps=1e-12):
    p = "embeddings."
    x = sd[p + "word_embeddings.weight"][ids] + sd[p + "token_type_embeddings.weight"][0]
    return _ln(sd, p + "LayerNorm", x, eps)                            # no position embedding (:190-208)


def _rope(table_rows, x):
    """x [B,H,S,64]; table_rows [S,64] = [sin(32) | cos(32)]  (xropebert.py:362-377)."""
    sin, cos = table_rows[None, None].chunk(2, dim=-1)
    sin_pos = torch.stack([sin, sin], dim=-1).reshape(1, 1, x.shape[2], -1)
    cos_pos = torch.stack([cos, cos], dim=-1).reshape(1, 1, x.shape[2], -1)
    rot = torch.stack([-x[..., 1::2], x[..., ::2]], dim=-1).reshape_as(x)
    return x * cos_pos + rot * sin_pos


def _sampler_attn(sd, p, x, kv, mask, heads, rope_q, rope_k):
    b, s, d = x.shape
    hd = d // heads

    def split(t):
        return t.view(b, -1, heads, hd).permute(0, 2, 1, 3)

    q = split(_lin(sd, p + "self.query", x))
    k = split(_lin(sd, p + "self.key", kv))
    v = split(_lin(sd, p + "self.value", kv))
    q = _rope(rope_q, q)
    k = _rope(rope_k, k)
    scores = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(hd)
    scores = scores + mask
    probs = F.softmax(scores, dim=-1)
    ctx = torch.matmul(probs, v).permute(0, 2, 1, 3).contiguous().view(b, s, d)
    return _ln(sd, p + "output.LayerNorm", _lin(sd, p + "output.dense", ctx) + x, 1e-12)


def sampler_forward(sd, of, of_mask, question_ids, question_mask, mode="fusion", layers=12, fusion_layer=6,
                    heads=12, eps=1e-12):
    """-> (sequence_output [B,Lof+2,768], logits [B,Lof,2])  (xropebert.py:1048-1169, return_dict falsy)."""
    x = of_embedding(sd, of, of_mask)
    txt = text_embedding(sd, question_ids, eps)
    self_mask = (1.0 - of_mask[:, None, None, :].to(x.dtype)) * -10000.0
    cross_mask = (1.0 - question_mask[:, None, None, :].to(x.dtype)) * torch.finfo(x.dtype).min
    rope_f = sd["encoder.embed_positions.weight"][: x.shape[1]]
    rope_t = sd["encoder.c_embed_positions.weight"][: txt.shape[1]]
    if mode in ("vision", "text"):
        lo, hi = 0, fusion_layer
    elif mode == "fusion":
        lo, hi = fusion_layer, layers
    elif mode == "multi_modal":
        lo, hi = 0, layers
    else:
        raise ValueError(f"INVALID MODE: {mode}")
    for i in range(lo, hi):
        p = f"encoder.layer.{i}."
        x = _sampler_attn(sd, p + "attention.", x, x, self_mask, heads, rope_f, rope_f)
        if i >= fusion_layer:
            x = _sampler_attn(sd, p + "crossattention.", x, txt, cross_mask, heads, rope_f, rope_t)
        h = _gelu(_lin(sd, p + "intermediate.dense", x))
        x = _ln(sd, p + "output.LayerNorm", _lin(sd, p + "output.dense", h) + x, eps)
    logits = _lin(sd, "mrc_head", x[:, 1:-1])
    return x, logits


def of_embedding_reassociated(sd, of, of_mask, patch=16):
    """The exact reassociation the CUDA path uses: weighted patch reduction
    z[f,(c,ky,kx)] = sum_p fc_w[p] * x[f,c,py*16+ky,px*16+kx], then ONE [.,512]x[512,768] GEMM.
    Kept here so tests can show it agrees with `of_embedding` to fp32 rounding."""
    b, l, c, h, w = of.shape
    p = "temporal_embeddings."
    g = h // patch
    fcw = sd[p + "fc.weight"].view(g, g)
    x = of.view(b * l, c, g, patch, g, patch)
    z = torch.einsum("fcykxj,yx->fckj", x, fcw).reshape(b * l, -1)                     # [B*L, 512]
    wc = sd[p + "projection.weight"].reshape(sd[p + "projection.weight"].shape[0], -1)   # [768, 512]
    y = z @ wc.t() + sd[p + "projection.bias"] * fcw.sum() + sd[p + "fc.bias"]
    y = y.view(b, l, -1)
    d = y.shape[-1]
    y = torch.cat([sd[p + "bos"].expand(b, 1, -1), y, torch.zeros(b, 1, d, device=y.device)], dim=1)
    ends = of_mask.sum(dim=1) - 1
    y[torch.arange(b), ends] = sd[p + "eos"]
    y = y + sd[p + "frame_pos_embed.weight"][: l + 2].unsqueeze(0)
    return F.layer_norm(y, (d,), sd[p + "ln.weight"], sd[p + "ln.bias"], 1e-5)
