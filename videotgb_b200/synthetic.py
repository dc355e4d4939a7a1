"""Seeded synthetic weights and inputs for the VideoTGB hot path.

There is no network for checkpoints, so tests, the golden-vector generator and
``bench.py`` all draw weights from here.  Key names and shapes follow the
reference state-dict contract (dumped from
``src/models/components/xblip2.py``, ``xinstructblip.py`` and ``xropebert.py``).

Initialisation deliberately differs from the reference's own ``_init_weights``
(``xblip2.py:281-285`` uses ``initializer_range = 1e-10`` for a stand-alone
vision model, which makes every parity test vacuous): every tensor is drawn
N(0, 0.02) -- LayerNorm gains 1 + N(0, 0.02) -- so that a wrong bias, a skipped
LayerNorm or a transposed weight is visible in the output.

Each tensor has its own generator seeded from ``crc32(name) ^ seed`` so that a
2-layer model shares its tensors with the 39-layer one.
"""
from __future__ import annotations

import zlib
from dataclasses import dataclass, field

import torch


@dataclass
class VisionCfg:
    hidden: int = 1408
    heads: int = 16
    layers: int = 39
    mlp: int = 6144
    image: int = 224
    patch: int = 14
    eps: float = 1e-6

    @property
    def tokens(self) -> int:
        return (self.image // self.patch) ** 2 + 1


@dataclass
class QFormerCfg:
    hidden: int = 768
    heads: int = 12
    layers: int = 12
    mlp: int = 3072
    enc_hidden: int = 1408
    cross_freq: int = 2
    n_query: int = 32
    eps: float = 1e-12
    vocab: int = 30522
    max_pos: int = 512
    instruct: bool = False


@dataclass
class SamplerCfg:
    hidden: int = 768
    heads: int = 12
    layers: int = 12
    fusion_layer: int = 6
    mlp: int = 3072
    vocab: int = 30522
    max_pos: int = 512
    eps: float = 1e-12
    image: int = 224
    patch: int = 16


def _draw(name: str, shape, seed: int, std: float = 0.02, mean: float = 0.0,
          device="cpu") -> torch.Tensor:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)
    t = torch.randn(tuple(shape), generator=g, dtype=torch.float32) * std + mean
    return t.to(device)


class _SD(dict):
    def __init__(self, prefix: str, seed: int, device):
        super().__init__()
        self._p, self._s, self._d = prefix, seed, device

    def w(self, name, *shape, std=0.02, mean=0.0):
        self[name] = _draw(self._p + name, shape, self._s, std, mean, self._d)

    def linear(self, name, out_f, in_f):
        self.w(name + ".weight", out_f, in_f)
        self.w(name + ".bias", out_f)

    def ln(self, name, n):
        self.w(name + ".weight", n, mean=1.0)
        self.w(name + ".bias", n)


def vision_state_dict(cfg: VisionCfg = VisionCfg(), seed: int = 1234, device="cpu") -> dict:
    """`Blip2VisionModel` / `InstructBlipVisionModel` keys (xblip2.py:89-117,120-263,500-563)."""
    sd = _SD("vision.", seed, device)
    h = cfg.hidden
    sd.w("embeddings.class_embedding", 1, 1, h)
    sd.w("embeddings.position_embedding", 1, cfg.tokens, h)
    sd.w("embeddings.patch_embedding.weight", h, 3, cfg.patch, cfg.patch)
    sd.w("embeddings.patch_embedding.bias", h)
    for i in range(cfg.layers):
        p = f"encoder.layers.{i}."
        sd.ln(p + "layer_norm1", h)
        sd.w(p + "self_attn.qkv.weight", 3 * h, h)
        b = _draw("vision." + p + "self_attn.qkv.bias", (3 * h,), seed, device=device)
        b[h:2 * h] = 0.0          # reference: cat(q_bias, zeros, v_bias)  (xblip2.py:140-149)
        sd[p + "self_attn.qkv.bias"] = b
        sd.linear(p + "self_attn.projection", h, h)
        sd.ln(p + "layer_norm2", h)
        sd.linear(p + "mlp.fc1", cfg.mlp, h)
        sd.linear(p + "mlp.fc2", h, cfg.mlp)
    sd.ln("post_layernorm", h)
    return dict(sd)


def qformer_state_dict(cfg: QFormerCfg = QFormerCfg(), seed: int = 1234, device="cpu") -> dict:
    """`Blip2QFormerModel` (xblip2.py:566-1174) / `InstructBlipQFormerModel` (xinstructblip.py:565-1242) keys."""
    sd = _SD("qformer.", seed, device)
    h = cfg.hidden
    if cfg.instruct:
        sd.w("embeddings.word_embeddings.weight", cfg.vocab, h)
        sd.w("embeddings.position_embeddings.weight", cfg.max_pos, h)
        sd.ln("embeddings.layernorm", h)
    else:
        sd.ln("layernorm", h)
    for i in range(cfg.layers):
        p = f"encoder.layer.{i}."
        for nm in ("query", "key", "value"):
            sd.linear(p + f"attention.attention.{nm}", h, h)
        sd.linear(p + "attention.output.dense", h, h)
        sd.ln(p + "attention.output.LayerNorm", h)
        if i % cfg.cross_freq == 0:
            sd.linear(p + "crossattention.attention.query", h, h)
            sd.linear(p + "crossattention.attention.key", h, cfg.enc_hidden)
            sd.linear(p + "crossattention.attention.value", h, cfg.enc_hidden)
            sd.linear(p + "crossattention.output.dense", h, h)
            sd.ln(p + "crossattention.output.LayerNorm", h)
        sd.linear(p + "intermediate_query.dense", cfg.mlp, h)
        sd.linear(p + "output_query.dense", h, cfg.mlp)
        sd.ln(p + "output_query.LayerNorm", h)
        if cfg.instruct:
            sd.linear(p + "intermediate.dense", cfg.mlp, h)
            sd.linear(p + "output.dense", h, cfg.mlp)
            sd.ln(p + "output.LayerNorm", h)
    return dict(sd)


def sampler_state_dict(cfg: SamplerCfg = SamplerCfg(), seed: int = 1234, device="cpu") -> dict:
    """`RopeBertModel` keys (xropebert.py:66-129,175-208,586-593,929-960)."""
    sd = _SD("sampler.", seed, device)
    h = cfg.hidden
    npatch = (cfg.image // cfg.patch) ** 2
    sd.w("embeddings.word_embeddings.weight", cfg.vocab, h)
    sd.w("embeddings.token_type_embeddings.weight", 2, h)
    sd.w("embeddings.position_embeddings.weight", cfg.max_pos, h)      # present, unused in forward
    sd.ln("embeddings.LayerNorm", h)
    sd.w("temporal_embeddings.bos", h)
    sd.w("temporal_embeddings.eos", h)
    sd["temporal_embeddings.position_ids"] = torch.arange(cfg.max_pos, device=device).expand(1, -1).clone()
    sd.w("temporal_embeddings.projection.weight", h, 2, cfg.patch, cfg.patch)
    sd.w("temporal_embeddings.projection.bias", h)
    sd.w("temporal_embeddings.fc.weight", 1, npatch, std=0.05)
    sd.w("temporal_embeddings.fc.bias", 1)
    sd.w("temporal_embeddings.frame_pos_embed.weight", cfg.max_pos, h)
    sd.ln("temporal_embeddings.ln", h)
    # "RoPE tables" are opaque frozen weights in real checkpoints.  Draw them
    # O(1) like real sin/cos so the rotation is not a near-no-op.
    hd = h // cfg.heads
    sd.w("encoder.embed_positions.weight", cfg.max_pos, hd, std=0.6)
    sd.w("encoder.c_embed_positions.weight", cfg.max_pos, hd, std=0.6)
    for i in range(cfg.layers):
        p = f"encoder.layer.{i}."
        for blk in (["attention"] + (["crossattention"] if i >= cfg.fusion_layer else [])):
            for nm in ("query", "key", "value"):
                sd.linear(p + f"{blk}.self.{nm}", h, h)
            sd.linear(p + f"{blk}.output.dense", h, h)
            sd.ln(p + f"{blk}.output.LayerNorm", h)
        sd.linear(p + "intermediate.dense", cfg.mlp, h)
        sd.linear(p + "output.dense", h, cfg.mlp)
        sd.ln(p + "output.LayerNorm", h)
    sd.linear("mrc_head", 2, h)
    return dict(sd)


def head_state_dict(llm_hidden: int = 2048, q_hidden: int = 768, n_query: int = 32, seed: int = 1234,
                    device="cpu") -> dict:
    """`query_tokens`, `language_projection`, `temporal_projection` (xblip2.py:1548-1552)."""
    sd = _SD("head.", seed, device)
    sd.w("query_tokens", 1, n_query, q_hidden)
    sd.linear("language_projection", llm_hidden, q_hidden)
    sd.linear("temporal_projection", llm_hidden, q_hidden)
    return dict(sd)


def synthetic_frames_u8(n: int, size: int = 224, seed: int = 1234) -> torch.Tensor:
    """u8 THWC frames ~ U{0..255} (SURVEY.md section 8d cfg 1)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (n, size, size, 3), generator=g, dtype=torch.uint8)


def synthetic_flow(b: int, lof: int, size: int = 224, seed: int = 1234) -> torch.Tensor:
    """Flow ~ N(0,1) scaled so max |uv| = 1 per video (mimics `normalize_flow`, videoinstruct_dataset.py:212-226)."""
    g = torch.Generator().manual_seed(seed + 1)
    of = torch.randn(b, lof, 2, size, size, generator=g)
    mag = of.pow(2).sum(2).sqrt().amax(dim=(1, 2, 3), keepdim=True).unsqueeze(2)
    return of / mag


def raft_state_dict(seed: int = 1234, device="cpu") -> dict:
    """`RAFT` keys (src/models/components/xraft.py:52-72; raft_utils/extractor.py, update.py): fnet (instance norm: no
    parameters), cnet (batch norm with non-trivial running statistics), update_block.  Conv weights ~ N(0, 2/fan_out)
    (the reference's kaiming_normal_ init) so that activations stay O(1) through 20 recurrent iterations."""
    sd = _SD("raft.", seed, device)

    def conv(name, cout, cin, kh, kw):
        sd.w(name + ".weight", cout, cin, kh, kw, std=(2.0 / (cout * kh * kw)) ** 0.5)
        sd.w(name + ".bias", cout, std=0.05)

    def bn(name, c):
        sd.w(name + ".weight", c, std=0.1, mean=1.0)
        sd.w(name + ".bias", c, std=0.1)
        sd.w(name + ".running_mean", c, std=0.1)
        sd[name + ".running_var"] = _draw("raft." + name + ".running_var", (c,), seed, 0.2, 0.0, device).abs() + 0.6
        sd[name + ".num_batches_tracked"] = torch.zeros((), dtype=torch.int64, device=device)

    for enc, out_dim, batch_norm in (("fnet.", 256, False), ("cnet.", 256, True)):
        conv(enc + "conv1", 64, 3, 7, 7)
        if batch_norm:
            bn(enc + "norm1", 64)
        cin = 64
        for li, dim in ((1, 64), (2, 96), (3, 128)):
            for bi in (0, 1):
                p = f"{enc}layer{li}.{bi}."
                conv(p + "conv1", dim, cin if bi == 0 else dim, 3, 3)
                conv(p + "conv2", dim, dim, 3, 3)
                if batch_norm:
                    bn(p + "norm1", dim)
                    bn(p + "norm2", dim)
                if bi == 0 and li > 1:
                    conv(p + "downsample.0", dim, cin, 1, 1)
                    if batch_norm:
                        bn(p + "norm3", dim)
                        for k in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked"):
                            sd[p + "downsample.1." + k] = sd[p + "norm3." + k]       # the same module under two names
            cin = dim
        conv(enc + "conv2", out_dim, 128, 1, 1)
    u = "update_block."
    conv(u + "encoder.convc1", 256, 324, 1, 1)
    conv(u + "encoder.convc2", 192, 256, 3, 3)
    conv(u + "encoder.convf1", 128, 2, 7, 7)
    conv(u + "encoder.convf2", 64, 128, 3, 3)
    conv(u + "encoder.conv", 126, 256, 3, 3)
    for g in "zrq":
        conv(u + f"gru.conv{g}1", 128, 384, 1, 5)
        conv(u + f"gru.conv{g}2", 128, 384, 5, 1)
    conv(u + "flow_head.conv1", 256, 128, 3, 3)
    conv(u + "flow_head.conv2", 2, 256, 3, 3)
    # a random-init update operator is not contractive (a trained one is): damp the flow head so that 20 iterations stay
    # in a regime where rounding differences do not grow chaotically (|flow| ~ 40 px on a 128x160 pair)
    sd[u + "flow_head.conv2.weight"] = sd[u + "flow_head.conv2.weight"] * 0.01
    sd[u + "flow_head.conv2.bias"] = sd[u + "flow_head.conv2.bias"] * 0.01
    conv(u + "mask.0", 256, 128, 3, 3)
    conv(u + "mask.2", 576, 256, 1, 1)
    return dict(sd)


def sinusoidal_table(n_pos: int, dim: int) -> torch.Tensor:
    """The rotary table the reference class constructs before `_init_weights` overwrites it with N(0, 0.02) draws
    (BertSinusoidalPositionalEmbedding, xropebert.py:149-164): sin of pos / 10000^(2j/dim) in the first half of a row,
    cos in the second half (not interleaved).  float64 arithmetic rounded to float32, as numpy -> FloatTensor does."""
    pos = torch.arange(n_pos, dtype=torch.float64)[:, None]
    j = torch.arange(dim // 2, dtype=torch.float64)[None, :]
    angle = pos / torch.pow(torch.tensor(10000.0, dtype=torch.float64), 2.0 * j / dim)
    return torch.cat([torch.sin(angle), torch.cos(angle)], dim=1).to(torch.float32)
