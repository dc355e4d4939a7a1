"""The hot path end to end: sampler scoring -> span draw / selection / gather -> ViT-g -> Q-Former ->
language_projection -> LLM query tokens.

`VideoTGBEncoder` has the attribute layout the reference task modules use
(`self.temporal_encoder`, `self.model.vision_model`, `self.model.qformer`, `self.model.query_tokens`,
`self.model.language_projection`, `self.model.temporal_projection`; src/models/LSTP_SF_blip2_module.py:84-89,
eval/utils/model.py:33-37), so `load_state_dict(ckpt["state_dict"], strict=False)` of a reference checkpoint fills it,
and `encode(batch)` is the body of `eval_forward` (LSTP_SF_blip2_module.py:383-484) up to the tokens handed to the
unchanged HF language model.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional

import torch
from torch import nn

from . import ops, selection
from .containers import (Blip2ForConditionalGeneration, InstructBlipForConditionalGeneration,  # noqa: F401
                         LanguageProjection)
from .modeling import _cached_bf16
from .sampler import RopeBertModel


class VideoTGBEncoder(nn.Module):
    def __init__(self, vision_config=None, qformer_config=None, sampler_config=None, llm_hidden: int = 2048,
                 instruct: bool = False, with_flow_extractor: bool = False):
        super().__init__()
        self.instruct = instruct
        cfg = SimpleNamespace(vision_config=vision_config, qformer_config=qformer_config, num_query_tokens=32,
                              text_config=SimpleNamespace(hidden_size=llm_hidden), use_decoder_only_language_model=instruct)
        # `self.model` of the reference task modules, without the (unchanged, out-of-scope) language model
        self.model = (InstructBlipForConditionalGeneration if instruct else Blip2ForConditionalGeneration)(
            cfg, language_model=False)
        self.temporal_encoder = RopeBertModel(sampler_config)
        if with_flow_extractor:            # offline inference computes the flow online (eval/utils/model.py:37,75-84)
            from .raft import RAFT
            self.of_extractor = RAFT()
        self.token_exchange = None      # optional distributed.TokenAllGather: all-gather of the Q-Former tokens

    @classmethod
    def from_modules(cls, model, temporal_encoder, of_extractor=None) -> "VideoTGBEncoder":
        """Wrap an existing container (`Blip2ForConditionalGeneration` / `InstructBlipForConditionalGeneration`, e.g. from
        `from_pretrained`) and sampler -- the two attributes the reference task modules hold (`self.model`,
        `self.temporal_encoder`, LSTP_SF_blip2_module.py:84-89) -- without rebuilding them."""
        self = cls.__new__(cls)
        nn.Module.__init__(self)
        self.instruct = bool(getattr(model, "instruct", False))
        self.model, self.temporal_encoder = model, temporal_encoder
        if of_extractor is not None:
            self.of_extractor = of_extractor
        self.token_exchange = None
        return self

    @torch.no_grad()
    def eval_forward(self, batch, generate_configs=None, mode: str = "fusion", gumbel=None, generator=None):
        """The whole `eval_forward` of the reference task modules (LSTP_SF_blip2_module.py:382-506, LSTP_SF_module.py:
        456-580): hot path -> `cat` with the question embeddings -> the UNCHANGED language model's `generate` -> the LLaMA
        id fix-up.  Needs a container that owns a language model."""
        lm = self.model.language_model
        if lm is None:
            raise RuntimeError("eval_forward needs a language model; this container was built with language_model=False")
        tokens = self.encode(batch, mode=mode, variant="eval", assemble="concat", gumbel=gumbel, generator=generator)
        inputs_embeds, attention_mask = self.llm_inputs(tokens, batch, self.model.get_input_embeddings().weight)
        outputs = lm.generate(inputs_embeds=inputs_embeds, attention_mask=attention_mask, **(generate_configs or {}))
        return self.fix_generated_ids(outputs, self.model.config.text_config.architectures[0])

    # ---- stages -------------------------------------------------------------------------------------------
    @torch.no_grad()
    def extract_flow(self, flow_frames) -> torch.Tensor:
        """SURVEY.md section 8f N1, eval/utils/model.py:75-84: per video `flow_frame` [T,3,H,W] -> RAFT on consecutive
        pairs (padded to a multiple of 8), last flow repeated -> `of` [B,T,2,H,W] for `temporal_encoder`."""
        from .raft import InputPadder
        of = []
        for flow_frame in flow_frames:
            padder = InputPadder(flow_frame[0].shape)
            ff = padder.pad(flow_frame)
            feats = self.of_extractor(ff[:-1], ff[1:])
            of.append(torch.cat([feats, feats[-1].unsqueeze(0)], dim=0))
        return torch.stack(of)

    def score(self, batch, mode: str = "fusion"):
        """A1-A6: -> (of_feat, start_logits [B,L,1], end_logits [B,L,1])."""
        of_feat, of_logits = self.temporal_encoder(
            encoder_embeds=batch["of"], attention_mask=batch["of_mask"],
            encoder_hidden_states=batch["sampler_question"],
            encoder_attention_mask=batch.get("sampler_question_attention_mask"), mode=mode)
        start_logits, end_logits = of_logits.split(1, dim=-1)
        return of_feat, start_logits, end_logits

    def select(self, start_logits, end_logits, batch, pixel_values, nframe, variant="eval", gumbel=None, generator=None):
        """A7-A9: -> (sampled_pixel_values [B*nframe,3,H,W], indices [B,nframe])."""
        return selection.select_and_gather(start_logits, end_logits, batch["of_lengths"], pixel_values, nframe,
                                           variant=variant, gumbel=gumbel, generator=generator)

    def tokens(self, sampled_pixel_values: torch.Tensor, batch_size: int, batch=None, assemble: str = "concat",
               widths=None) -> torch.Tensor:
        """A10-A16: frames -> `language_model_inputs` ([B, nframe*32, H_llm] concat, or [B, 32, H_llm] mean-pooled)."""
        image_embeddings = self.model.vision_model(pixel_values=sampled_pixel_values, return_dict=True).last_hidden_state
        return self.assemble(self._query_output(image_embeddings, batch_size, batch), batch_size, assemble, widths)

    def _query_output(self, image_embeddings: torch.Tensor, batch_size: int, batch=None) -> torch.Tensor:
        """A14 / A15 as the task modules call the Q-Former (LSTP_SF_blip2_module.py:471-477, LSTP_SF_module.py:545-560)."""
        m = self.model
        nf = image_embeddings.shape[0]
        query_tokens = m.query_tokens.expand(nf, -1, -1)
        if self.instruct:
            nframe = nf // batch_size
            ids = torch.repeat_interleave(batch["qformer_text"], nframe, 0)
            tmask = torch.repeat_interleave(batch["qformer_text_attention_mask"], nframe, 0)
            qmask = torch.ones(query_tokens.shape[:-1], dtype=torch.long, device=ids.device)
            out = m.qformer(input_ids=ids, attention_mask=torch.cat([qmask, tmask], dim=1), query_embeds=query_tokens,
                            encoder_hidden_states=image_embeddings, encoder_attention_mask=None, return_dict=True)
            return out.last_hidden_state[:, : query_tokens.size(1), :]
        slot = None
        if self.token_exchange is not None:          # bf16 tokens go from the last LayerNorm straight into the NCCL buffer
            slot = self.token_exchange.slot(nf, query_tokens.shape[1], m.qformer.hidden, image_embeddings.device)
        return m.qformer(query_embeds=query_tokens, encoder_hidden_states=image_embeddings, encoder_attention_mask=None,
                         tokens_out_bf16=slot)[0]

    def assemble(self, query_output: torch.Tensor, batch_size: int, mode: str = "concat", widths=None) -> torch.Tensor:
        m = self.model
        if self.token_exchange is not None:
            self.token_exchange.exchange(query_output)      # NCCL all-gather on a side stream, overlaps the projection
        if mode == "concat":                       # LSTP_SF_blip2_module.py:480-481
            pre = _cached_bf16(query_output)      # the Q-Former's own bf16 rounding of these values: no re-cast
            y = m.language_projection(query_output if pre is None else pre)
            return y.reshape(batch_size, -1, y.shape[-1])
        if mode != "mean":
            raise ValueError(mode)
        nf = query_output.shape[0]                 # eval/utils/model.py:186-195, LSTP_Vicuna_IV_module.py:233-242
        q = query_output.to(torch.float32).contiguous()
        if widths is None or len(set(widths)) == 1:
            pooled = ops.group_mean(q, batch_size, nf // batch_size)
        else:
            parts, idx = [], 0
            for w in widths:
                parts.append(ops.group_mean(q[idx: idx + w].contiguous(), 1, w))
                idx += w
            pooled = torch.cat(parts, 0)
        return m.language_projection(pooled)

    def llm_inputs(self, language_model_inputs: torch.Tensor, batch, embedding_weight: torch.Tensor):
        """SURVEY.md section 8f N4, LSTP_SF_blip2_module.py:482-495: `(inputs_embeds, attention_mask)` for the unchanged
        language model's `generate`: visual tokens followed by the question's word embeddings
        (`embedding_weight` = `language_model.get_input_embeddings().weight`)."""
        if embedding_weight.dtype != torch.float32:
            raise ValueError("the hand-off kernel copies fp32 embeddings (the reference keeps the LLM embedding in fp32)")
        if int(batch["question"].min()) < 0 or int(batch["question"].max()) >= embedding_weight.shape[0]:
            raise IndexError("index out of range in self")       # what nn.Embedding raises
        return ops.llm_inputs(language_model_inputs.contiguous(), embedding_weight.contiguous(), batch["question"].contiguous(),
                              batch["question_attention_mask"].to(torch.int64).contiguous())

    @staticmethod
    def fix_generated_ids(outputs: torch.Tensor, architecture: str = "LLaMAForCausalLM") -> torch.Tensor:
        """LSTP_SF_blip2_module.py:503-504: `outputs[outputs==0] = 2` for LLaMA text configs (in place)."""
        return ops.replace_ids_(outputs, 0, 2) if architecture == "LLaMAForCausalLM" else outputs

    def tokens_from_u8(self, sampled_u8: torch.Tensor, batch_size: int, batch=None, assemble: str = "concat"):
        """Same as `tokens` for u8 THWC frames: normalise (A0) is fused into the patchify kernel."""
        image_embeddings = self.model.vision_model.forward_u8(sampled_u8)
        return self.assemble(self._query_output(image_embeddings, batch_size, batch), batch_size, assemble)

    # ---- the whole path -----------------------------------------------------------------------------------
    @torch.no_grad()
    def encode(self, batch, mode: str = "fusion", variant: str = "eval", assemble: str = "concat", gumbel=None,
               generator=None, return_aux: bool = False):
        """batch keys as the reference collate produces them (videoinstruct_dataset.py:147-192): frames [B*N,3,H,W]
        (or [B,N,3,H,W]), of, of_mask, of_lengths, sampler_question(+_attention_mask), nframe, and for InstructBLIP
        qformer_text(+_attention_mask)."""
        nframe = batch["nframe"]
        b = batch["of"].shape[0]
        px = batch["frames"]
        if px.dim() == 4:
            px = px.view(b, -1, *px.shape[1:])
        _, start_logits, end_logits = self.score(batch, mode)
        sampled, idx = self.select(start_logits, end_logits, batch, px, nframe, variant, gumbel, generator)
        out = self.tokens(sampled, b, batch, assemble)
        if return_aux:
            return out, SimpleNamespace(indices=idx, start_logits=start_logits, end_logits=end_logits)
        return out

    @torch.no_grad()
    def encode_sharded_u8(self, batch, group=None, mode: str = "fusion", variant: str = "eval", gumbel=None, generator=None,
                          return_aux: bool = False):
        """BASELINE config 4: ONE long video whose candidates are sharded by frame over `group` (rank r holds frames and
        flow [r*T_local, (r+1)*T_local) in `frames_u8` [1,T_local,H,W,3] and `of` [1,T_local,2,H,W]; `of_mask`,
        `of_lengths`, the question and `nframe` describe the whole video).  Scoring is sharded by candidate, the
        selection is replicated, and the `nframe` selected frames are then re-distributed EVENLY (all-to-all of the u8
        frames) so that every rank encodes nframe/W of them.  Every rank returns the same `language_model_inputs`
        [1, nframe*32, H_llm] the unsharded `encode_u8` produces, bit for bit (replaces the
        process-level chunking of eval/scripts/run_qa_activitynet.sh:16-39)."""
        import torch.distributed as dist
        from . import distributed as D
        if self.instruct:
            raise NotImplementedError("frame-sharded long videos are wired for the BLIP-2 Q-Former (config 4)")
        if group is None:
            group = dist.group.WORLD
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        nframe = batch["nframe"]
        px = batch["frames_u8"]
        if px.shape[0] != 1:
            raise ValueError("encode_sharded_u8 handles one long video; shard whole videos with shard_videos() otherwise")
        t_local = px.shape[1]
        of_feat, of_logits = self.temporal_encoder(
            encoder_embeds=batch["of"], attention_mask=batch["of_mask"], encoder_hidden_states=batch["sampler_question"],
            encoder_attention_mask=batch.get("sampler_question_attention_mask"), mode=mode, flow_group=group)
        start_logits, end_logits = of_logits.split(1, dim=-1)
        if gumbel is None:                                  # identical noise everywhere: drawn on rank 0, broadcast
            lg = torch.cat([start_logits.reshape(1, -1), end_logits.reshape(1, -1)], 0)
            gumbel = [selection.gumbel_noise_like(lg, generator) for _ in range(2)]
            for g in gumbel:
                dist.broadcast(g, src=dist.get_global_rank(group, 0), group=group)
        cs, ce = selection.draw_spans(start_logits, end_logits, 2, 0.5, gumbel, None)
        idx = selection.select_indices(cs, ce, batch["of_lengths"], t_local * world, nframe, variant)
        plan = D.encode_plan(idx[0].tolist(), rank, world, t_local)     # one host sync per video (the reference syncs per comparison)
        counts = plan.counts
        m = self.model
        sampled = D.redistribute_selected_frames(px.contiguous(), plan, group)          # [k/W, H, W, 3] u8 per rank
        nq, hid = m.query_tokens.shape[1], m.qformer.hidden
        if sampled.shape[0]:
            image_embeddings = m.vision_model.forward_u8(sampled)
            query_tokens = m.query_tokens.expand(image_embeddings.shape[0], -1, -1)
            qout = m.qformer(query_embeds=query_tokens, encoder_hidden_states=image_embeddings)[0]
            tokens_local = _cached_bf16(qout)
            if tokens_local is None:
                tokens_local = ops.cast_bf16(qout.to(torch.float32).contiguous().view(-1, hid)).view(-1, nq, hid)
        else:
            tokens_local = torch.zeros((0, nq, hid), dtype=torch.bfloat16, device=px.device)
        tokens = D.all_gather_encoded_tokens(tokens_local, counts, group)               # bf16 [k, 32, 768]
        y = m.language_projection(tokens)
        out = y.reshape(1, -1, y.shape[-1])
        if return_aux:
            return out, SimpleNamespace(indices=idx, start_logits=start_logits, end_logits=end_logits, counts=counts)
        return out

    @torch.no_grad()
    def encode_u8(self, batch, mode: str = "fusion", variant: str = "eval", assemble: str = "concat", gumbel=None,
                  generator=None, return_aux: bool = False):
        """Same path starting from DECODED u8 frames `frames_u8` [B,N,H,W,3] (what the dataloader holds before the
        reference's CPU-side ToTensorVideo/NormalizeVideo).  Normalisation is a per-pixel map, so it commutes with the
        gather: only the `nframe` selected frames per video are normalised, fused into the patchify kernel -- the fp32
        [B,N,3,H,W] candidate tensor of the reference is never materialised.  Results equal `encode` on
        `VideoNormalizer()(frames_u8)`."""
        nframe = batch["nframe"]
        px = batch["frames_u8"]
        b = px.shape[0]
        _, start_logits, end_logits = self.score(batch, mode)
        cs, ce = selection.draw_spans(start_logits, end_logits, 2, 0.5, gumbel, generator)
        idx = selection.select_indices(cs, ce, batch["of_lengths"], px.shape[1], nframe, variant)
        sampled = ops.gather_frames(px.contiguous(), idx)
        out = self.tokens_from_u8(sampled, b, batch, assemble)
        if return_aux:
            return out, SimpleNamespace(indices=idx, start_logits=start_logits, end_logits=end_logits)
        return out
