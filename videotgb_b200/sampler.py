"""Temporal sampler: drop-in for the reference `RopeBertModel` (src/models/components/xropebert.py:929-1178).

Call contract (LSTP_SF_blip2_module.py:408-414, eval/utils/model.py:92-100):

    seq_out, logits = temporal_encoder(encoder_embeds=of[B,Lof,2,H,W], attention_mask=of_mask[B,Lof+2],
                                       encoder_hidden_states=question_ids[B,Tq], encoder_attention_mask=[B,Tq],
                                       mode="fusion" | "multi_modal")

returns the 2-tuple `(sequence_output [B,Lof+2,768], logits [B,Lof,2])` (the reference's `return_dict` falsy path,
xropebert.py:1167-1169); with `return_dict=True` the logits ride in `pooler_output` as in the reference (:1171-1173).

Stages (all on the sm_100a kernels, no fallback):
  A1  flow -> weighted patch reduction (HBM-bound, one pass over the flow) -> [T,512]x[512,768] tcgen05 GEMM ->
      bos/eos/pos/LayerNorm(1e-5)
  A2  question ids -> word + token_type[0] -> LayerNorm
  A4/A5  6 or 12 post-LN BERT layers; RoPE tables are opaque checkpoint weights (`encoder.embed_positions.weight`,
      `encoder.c_embed_positions.weight`)
  A6  mrc_head in fp32
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Optional

import torch
from torch import nn

from . import ops
from .modeling import _Conv, _Embedding, _Holder, _KernelModule, _LayerNorm, _Linear, _cfg, _model_output


class RopeBertModel(_KernelModule):
    def __init__(self, config=None, add_pooling_layer: bool = True):
        super().__init__()
        self.config = config if config is not None else SimpleNamespace()
        d = self.hidden = _cfg(config, "hidden_size", 768)
        self.heads = _cfg(config, "num_attention_heads", 12)
        self.layers_n = _cfg(config, "num_hidden_layers", 12)
        self.fusion_layer = _cfg(config, "fusion_layer", 6)
        self.mlp = _cfg(config, "intermediate_size", 3072)
        self.enc_width = _cfg(config, "encoder_width", d)
        self.eps = _cfg(config, "layer_norm_eps", 1e-12)
        vocab = _cfg(config, "vocab_size", 30522)
        max_pos = _cfg(config, "max_position_embeddings", 512)
        types = _cfg(config, "type_vocab_size", 2)
        self.head_dim = d // self.heads
        self.graph_max_tokens = 8192       # layer stacks over at most this many rows replay a CUDA graph (0 disables)
        self.flow_patch, self.flow_size = 16, 224
        if d % self.heads != 0:
            raise ValueError(f"The hidden size ({d}) is not a multiple of the number of attention heads ({self.heads})")
        self.embeddings = _Holder()
        self.embeddings.word_embeddings = _Embedding(vocab, d)
        self.embeddings.token_type_embeddings = _Embedding(types, d)
        self.embeddings.position_embeddings = _Embedding(max_pos, d)      # present in checkpoints, unused in forward
        self.embeddings.LayerNorm = _LayerNorm(d, self.eps)
        te = self.temporal_embeddings = _Holder()
        te.projection = _Conv(2, d, self.flow_patch)
        te.fc = _Linear((self.flow_size // self.flow_patch) ** 2, 1)
        te.bos = nn.Parameter(torch.empty(d).normal_(0.0, 0.02))
        te.eos = nn.Parameter(torch.empty(d).normal_(0.0, 0.02))
        te.register_buffer("position_ids", torch.arange(max_pos).expand(1, -1).clone())
        te.frame_pos_embed = _Embedding(max_pos, d)
        te.ln = _LayerNorm(d, 1e-5)                                        # nn.LayerNorm default eps (xropebert.py:91)
        self.encoder = _Holder()
        self.encoder.embed_positions = _Embedding(max_pos, self.head_dim)
        self.encoder.c_embed_positions = _Embedding(max_pos, self.head_dim)
        self.encoder.layer = nn.ModuleList()
        for i in range(self.layers_n):
            lyr = _Holder()
            lyr.attention = self._attn(d, d)
            if i >= self.fusion_layer:
                lyr.crossattention = self._attn(d, self.enc_width)
            lyr.intermediate = _Holder()
            lyr.intermediate.dense = _Linear(d, self.mlp)
            lyr.output = _Holder()
            lyr.output.dense = _Linear(self.mlp, d)
            lyr.output.LayerNorm = _LayerNorm(d, self.eps)
            self.encoder.layer.append(lyr)
        self.mrc_head = _Linear(d, 2)

    def _attn(self, d, kv_in):
        blk = _Holder()
        setattr(blk, "self", _Holder())
        s = getattr(blk, "self")
        s.query, s.key, s.value = _Linear(d, d), _Linear(kv_in, d), _Linear(kv_in, d)
        blk.output = _Holder()
        blk.output.dense = _Linear(d, d)
        blk.output.LayerNorm = _LayerNorm(d, self.eps)
        return blk

    def get_input_embeddings(self):
        return self.embeddings.word_embeddings

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, config=None, **unused):
        """`RopeBertModel.from_pretrained(sampler_name_or_path, config=BertConfig(fusion_layer=6, encoder_width=768))`
        (LSTP_SF_blip2_module.py:88-89).  The reference initialises from a plain BERT checkpoint (`bert-base-uncased`):
        HF strips the `bert.` base-model prefix (xropebert.py:805) and renames legacy `LayerNorm.gamma/beta`; everything
        the checkpoint does not hold -- temporal_embeddings, cross-attention, the rotary tables, mrc_head -- keeps its
        fresh initialisation (N(0, 0.02) weights, xropebert.py:821-835).  A VideoTGB checkpoint (keys already in this
        model's naming, optionally under `temporal_encoder.`) loads the same way."""
        from .containers import _report, read_checkpoint
        if config is None:
            import transformers
            config = transformers.BertConfig.from_pretrained(pretrained_model_name_or_path)
        model = cls(config)
        own = model.state_dict()
        mapped, unexpected = {}, []
        for k, v in read_checkpoint(pretrained_model_name_or_path).items():
            k2 = k
            for prefix in ("temporal_encoder.", "bert."):
                if k2.startswith(prefix):
                    k2 = k2[len(prefix):]
            k2 = k2.replace("LayerNorm.gamma", "LayerNorm.weight").replace("LayerNorm.beta", "LayerNorm.bias")
            if k2 in own and tuple(own[k2].shape) == tuple(v.shape):
                mapped[k2] = v.float() if v.is_floating_point() else v
            elif not k2.endswith("position_ids"):
                unexpected.append(k)
        res = model.load_state_dict(mapped, strict=False)
        model.loaded_keys = sorted(mapped)
        # a plain BERT checkpoint leaves the sampler-specific parameters at their initialisation, as in the reference
        _report([], [k for k in unexpected if not k.startswith(("cls.", "bert.pooler.", "model.", "of_extractor."))],
                f"RopeBertModel.from_pretrained({pretrained_model_name_or_path})")
        model.missing_keys = list(res.missing_keys)
        model.eval()
        return model

    def _pack(self):
        if self._packed is not None:
            return self._packed
        te = self.temporal_embeddings
        d = self.hidden
        fcw = self._f32(te.fc.weight.reshape(-1))
        pk = SimpleNamespace(
            fcw=fcw,
            conv_w=self._bf16(te.projection.weight.reshape(d, -1)),                       # [768, 512]
            conv_b=(te.projection.bias.detach().float() * fcw.sum() + te.fc.bias.detach().float()).contiguous(),
            bos=self._f32(te.bos), eos=self._f32(te.eos), fpos=self._f32(te.frame_pos_embed.weight),
            te_ln=(self._f32(te.ln.weight), self._f32(te.ln.bias)),
            word=self._f32(self.embeddings.word_embeddings.weight),
            type0=self._f32(self.embeddings.token_type_embeddings.weight[0]),
            emb_ln=(self._f32(self.embeddings.LayerNorm.weight), self._f32(self.embeddings.LayerNorm.bias)),
            rope_f=self._f32(self.encoder.embed_positions.weight), rope_t=self._f32(self.encoder.c_embed_positions.weight),
            mrc_w=self._f32(self.mrc_head.weight), mrc_b=self._f32(self.mrc_head.bias), layers=[])
        for lyr in self.encoder.layer:
            a = getattr(lyr.attention, "self")
            lp = SimpleNamespace(
                qkv_w=self._bf16(torch.cat([a.query.weight, a.key.weight, a.value.weight], 0)),
                qkv_b=self._f32(torch.cat([a.query.bias, a.key.bias, a.value.bias], 0)),
                ao_w=self._bf16(lyr.attention.output.dense.weight), ao_b=self._f32(lyr.attention.output.dense.bias),
                ao_ln=(self._f32(lyr.attention.output.LayerNorm.weight), self._f32(lyr.attention.output.LayerNorm.bias)),
                cross=None,
                i_w=self._bf16(lyr.intermediate.dense.weight), i_b=self._f32(lyr.intermediate.dense.bias),
                o_w=self._bf16(lyr.output.dense.weight), o_b=self._f32(lyr.output.dense.bias),
                o_ln=(self._f32(lyr.output.LayerNorm.weight), self._f32(lyr.output.LayerNorm.bias)))
            if hasattr(lyr, "crossattention"):
                c = getattr(lyr.crossattention, "self")
                lp.cross = SimpleNamespace(
                    q_w=self._bf16(c.query.weight), q_b=self._f32(c.query.bias),
                    kv_w=self._bf16(torch.cat([c.key.weight, c.value.weight], 0)),
                    kv_b=self._f32(torch.cat([c.key.bias, c.value.bias], 0)),
                    o_w=self._bf16(lyr.crossattention.output.dense.weight), o_b=self._f32(lyr.crossattention.output.dense.bias),
                    o_ln=(self._f32(lyr.crossattention.output.LayerNorm.weight),
                          self._f32(lyr.crossattention.output.LayerNorm.bias)))
            pk.layers.append(lp)
        self._packed = pk
        return pk

    # A1
    def embed_flow(self, of: torch.Tensor, of_mask: torch.Tensor, flow_group=None, ends: Optional[torch.Tensor] = None):
        """-> (f32, bf16) [B, Lof+2, 768]  (TemporalOFEmbedding.forward, xropebert.py:103-129).

        `flow_group`: the flow is sharded by frame over this process group (`of` holds this rank's contiguous slice,
        `of_mask` covers the whole video); the per-frame embeddings are all-gathered before bos/eos/LayerNorm."""
        pk = self._pack()
        b, l, c, h, w = of.shape
        if c != 2 or h != w:
            raise ValueError("flow must be [B, L, 2, S, S]")
        l_total = of_mask.shape[1] - 2 if flow_group is not None else l
        if l_total + 2 > pk.fpos.shape[0]:
            raise ValueError(f"{l_total} flow frames need {l_total + 2} positions > max_position_embeddings={pk.fpos.shape[0]} "
                             "(the reference indexes out of range here; build the config with a larger table)")
        z = ops.of_patch_reduce(of.to(torch.float32).contiguous().view(b * l, c, h, w), pk.fcw, self.flow_patch)
        y = ops.gemm(z, pk.conv_w, pk.conv_b, ops.EPI_F32).view(b, l, -1)
        if flow_group is not None:
            from .distributed import all_gather_frame_rows
            y = all_gather_frame_rows(y, flow_group)
            if y.shape[1] != l_total:
                raise ValueError(f"sharded flow covers {y.shape[1]} frames but of_mask describes {l_total}")
        if ends is None:
            ends = ops.additive_mask(of_mask, 0.0, want_out=False, want_last_index=True)[1]
        return ops.sampler_assemble(y.contiguous(), pk.bos, pk.eos, pk.fpos, ends, pk.te_ln[0], pk.te_ln[1], 1e-5)

    def forward(self, input_ids=None, attention_mask=None, token_type_ids=None, head_mask=None, inputs_embeds=None,
                encoder_embeds: Optional[torch.Tensor] = None, encoder_hidden_states: Optional[torch.Tensor] = None,
                encoder_attention_mask: Optional[torch.Tensor] = None, past_key_values=None, use_cache=None,
                output_attentions=None, output_hidden_states=None, return_dict=None, is_decoder=False,
                mode: str = "multi_modal", flow_group=None):
        if input_ids is not None and inputs_embeds is not None:
            raise ValueError("You cannot specify both input_ids and inputs_embeds at the same time")
        if encoder_embeds is None:
            if input_ids is None and inputs_embeds is None:
                raise ValueError("You have to specify either input_ids or inputs_embeds or encoder_embeds")
            raise NotImplementedError("only the encoder_embeds (optical-flow) entry of RopeBertModel is on the hot path")
        if mode in ("vision", "text"):
            lo, hi = 0, self.fusion_layer
        elif mode == "fusion":
            lo, hi = self.fusion_layer, self.layers_n
        elif mode == "multi_modal":
            lo, hi = 0, self.layers_n
        else:
            raise ValueError(f"INVALID MODE: {mode}")
        if encoder_hidden_states is None:
            raise ValueError("encoder_hidden_states (question token ids) are required with encoder_embeds")
        self._device_check(encoder_embeds)
        pk = self._pack()
        d, h, hd = self.hidden, self.heads, self.head_dim
        b, l = encoder_embeds.shape[:2]
        if flow_group is not None:
            if attention_mask is None:
                raise ValueError("a frame-sharded flow needs the attention_mask of the whole video")
            l = attention_mask.shape[1] - 2
        s = l + 2
        if attention_mask is None:
            attention_mask = torch.ones((b, s), dtype=torch.long, device=encoder_embeds.device)
        # A3: both additive masks and the eos position in two launches (the reference: ~10 elementwise statements)
        self_mask, ends = ops.additive_mask(attention_mask, -10000.0, want_last_index=True)
        x_f32, x_bf16 = self.embed_flow(encoder_embeds, attention_mask, flow_group, ends=ends)
        x_f32, x_bf16 = x_f32.view(-1, d), x_bf16.view(-1, d)
        ids = encoder_hidden_states.to(torch.int64).contiguous()
        tq = ids.shape[1]
        if tq > pk.rope_t.shape[0] or s > pk.rope_f.shape[0]:
            raise ValueError("sequence longer than the rotary position table")
        _, txt = ops.embed_layernorm(ids.view(-1), pk.word, pk.emb_ln[0], pk.emb_ln[1], self.eps, addvec=pk.type0)
        # no text mask = all ones = a zero additive mask = none at all (the kernels treat NULL as zeros)
        cross_mask = None if encoder_attention_mask is None else \
            ops.additive_mask(encoder_attention_mask, torch.finfo(torch.float32).min)[0]
        # The layer stack is ~17 launches per layer of a few microseconds each (M = B*(Lof+2) rows): launch-bound from
        # Python, so a shape seen for the second time is captured and replayed as one CUDA graph.
        have_cross_mask = cross_mask is not None
        inputs = [x_f32, x_bf16, txt, self_mask] + ([cross_mask] if have_cross_mask else [])

        def body(x_f32, x_bf16, txt, self_mask, cross_mask=None):
            return self._run_layers(pk, lo, hi, x_f32, x_bf16, txt, self_mask, cross_mask, b, s, tq)

        if b * s <= self.graph_max_tokens and self._use_graph(1):
            x_f32, logits_all = self._graphs.run(("sampler", lo, hi, b, s, tq, have_cross_mask), inputs, body, min_seen=2)
        else:
            x_f32, logits_all = body(*inputs)
        seq_out = x_f32.view(b, s, d)
        logits = logits_all.view(b, s, 2)[:, 1:-1]
        if not return_dict:
            return seq_out, logits
        return _model_output("BaseModelOutputWithPoolingAndCrossAttentions", last_hidden_state=seq_out,
                             pooler_output=logits, past_key_values=None, hidden_states=None, attentions=None,
                             cross_attentions=None)

    def _run_layers(self, pk, lo, hi, x_f32, x_bf16, txt, self_mask, cross_mask, b, s, tq):
        """Layers [lo, hi) + the mrc head on [B*s, 768] rows -> (sequence output f32 [B*s, 768], logits f32 [B*s, 2])."""
        d, h, hd = self.hidden, self.heads, self.head_dim
        scale = 1.0 / math.sqrt(hd)
        for i in range(lo, hi):
            lp = pk.layers[i]
            qkv = ops.gemm(x_bf16, lp.qkv_w, lp.qkv_b, ops.EPI_BF16)
            ops.rope_(qkv[:, :d], pk.rope_f, s, h, hd)
            ops.rope_(qkv[:, d:2 * d], pk.rope_f, s, h, hd)
            ctx = ops.attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], b, h, s, s, hd, scale, self_mask)
            acc = ops.gemm(ctx, lp.ao_w, lp.ao_b, ops.EPI_F32_RESID, resid=x_f32)
            x_bf16, x_f32 = ops.layernorm(acc, lp.ao_ln[0], lp.ao_ln[1], self.eps, want_bf16=True, want_f32=True)
            if lp.cross is not None:
                c = lp.cross
                q = ops.gemm(x_bf16, c.q_w, c.q_b, ops.EPI_BF16)
                ops.rope_(q, pk.rope_f, s, h, hd)
                kv = ops.gemm(txt, c.kv_w, c.kv_b, ops.EPI_BF16)
                ops.rope_(kv[:, :d], pk.rope_t, tq, h, hd)
                ctx = ops.attention(q, kv[:, :d], kv[:, d:], b, h, s, tq, hd, scale, cross_mask)
                acc = ops.gemm(ctx, c.o_w, c.o_b, ops.EPI_F32_RESID, resid=x_f32)
                x_bf16, x_f32 = ops.layernorm(acc, c.o_ln[0], c.o_ln[1], self.eps, want_bf16=True, want_f32=True)
            inter = ops.gemm(x_bf16, lp.i_w, lp.i_b, ops.EPI_BF16_GELU)
            acc = ops.gemm(inter, lp.o_w, lp.o_b, ops.EPI_F32_RESID, resid=x_f32)
            x_bf16, x_f32 = ops.layernorm(acc, lp.o_ln[0], lp.o_ln[1], self.eps, want_bf16=True, want_f32=True)
        return x_f32, ops.linear_small(x_f32, pk.mrc_w, pk.mrc_b)
