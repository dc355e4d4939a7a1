"""Span draw -> frame indices -> gather, on the device with no host round trip.

Replaces the inline block every reference task module carries (canonical: src/models/LSTP_SF_blip2_module.py:415-464):
two rounds of `F.gumbel_softmax(cat(start,end), tau=0.5, dim=1).argmax(dim=1)`, the per-video Python loop (span ->
frame range, set union, fallback, doubling, numpy-linspace midpoints -- each comparison on a CUDA scalar is a device
sync in the reference) and `index_select` into an fp32 zeros buffer.  Here it is three kernel launches.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import ops

VARIANTS = {
    # name: (mapping, inclusive)         reference site
    "eval": (ops.MAP_SCALED, False),     # LSTP_SF_blip2_module.py:446-447
    "train": (ops.MAP_SCALED, True),     # LSTP_SF_blip2_module.py:296-297
    "ratio": (ops.MAP_RATIO, False),     # LSTP_module.py:245, eval/utils/model.py:135
}


def gumbel_noise_like(logits: torch.Tensor, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """Exactly the draw `F.gumbel_softmax` makes: -log(Exp(1)) with the tensor's device RNG."""
    e = torch.empty_like(logits, memory_format=torch.legacy_contiguous_format)
    return -e.exponential_(generator=generator).log()


def draw_spans(start_logits: torch.Tensor, end_logits: torch.Tensor, top_k: int = 2, tau: float = 0.5,
               gumbel: Optional[Sequence[torch.Tensor]] = None, generator: Optional[torch.Generator] = None):
    """start/end logits [B,L,1] (or [B,L]) -> (cand_start, cand_end) int64 [top_k, B] on the device.

    `gumbel[i]` ([2B,L,1] or [2B,L]) overrides the i-th round's noise (parity tests feed the reference's draw).
    """
    b, l = start_logits.shape[:2]
    sl = start_logits.reshape(b, l).to(torch.float32).contiguous()          # no-ops for the sampler's fp32 logits
    el = end_logits.reshape(b, l).to(torch.float32).contiguous()
    cand = torch.empty((2, top_k, b), dtype=torch.int64, device=sl.device)
    for i in range(top_k):
        if gumbel is not None:
            ops.span_draw(sl, el, gumbel[i].reshape(2 * b, l).to(torch.float32).contiguous(), False, tau, cand[0, i], cand[1, i])
        else:
            # the reference's draw, `-torch.empty_like(logits).exponential_().log()`: the Exp(1) kernel stays torch's (same
            # generator stream as F.gumbel_softmax), -log and everything after it run inside the one kernel
            e = torch.empty((2 * b, l), dtype=torch.float32, device=sl.device).exponential_(generator=generator)
            ops.span_draw(sl, el, e, True, tau, cand[0, i], cand[1, i])
    return cand[0], cand[1]


def select_indices(cand_start: torch.Tensor, cand_end: torch.Tensor, of_lengths, num_frames: int, nframe: int,
                   variant: str = "eval", host_ints: bool = False) -> torch.Tensor:
    """-> int64 [B, nframe] sorted frame indices (device)."""
    mapping, inclusive = VARIANTS[variant]
    if not torch.is_tensor(of_lengths) or not of_lengths.is_cuda:
        # host-side lengths (the reference's batch["of_lengths"] is a Python list): same error as the reference's
        # `... / (video_length - 1)`.  A CUDA tensor is not synchronised on: the kernel treats a length below 2 as 2.
        host = [int(v) for v in (of_lengths.tolist() if torch.is_tensor(of_lengths) else of_lengths)]
        if min(host) < 2:
            raise ZeroDivisionError("of_lengths must be >= 2 (the reference divides by video_length-1)")
        of_lengths = torch.tensor(host, dtype=torch.int32, device=cand_start.device)
    return ops.spans_to_indices(cand_start.contiguous(), cand_end.contiguous(), of_lengths.to(torch.int32).contiguous(),
                                num_frames, nframe, mapping, inclusive, host_ints)


def uniform_indices(batch: int, num_frames: int, nframe: int, device) -> torch.Tensor:
    """The uniform-sampling baseline / fallback (`cand_index = list(range(num_frames))`, LSTP_blip2_module.py:254):
    an empty candidate set falls back to all frames inside the kernel."""
    z = torch.zeros((1, batch), dtype=torch.int64, device=device)
    lengths = torch.full((batch,), max(2, num_frames), dtype=torch.int32, device=device)
    # a reversed span (1, 0) maps to an empty range -> "all frames" -> linspace midpoints
    return ops.spans_to_indices(z + 1, z, lengths, num_frames, nframe, ops.MAP_SCALED, False, True)


def select_and_gather(start_logits, end_logits, of_lengths, pixel_values: torch.Tensor, nframe: int,
                      variant: str = "eval", top_k: int = 2, tau: float = 0.5, gumbel=None, generator=None):
    """pixel_values [B,N,3,H,W] -> (sampled [B*nframe,3,H,W] (source dtype; fp32 in the reference), indices [B,nframe])."""
    cs, ce = draw_spans(start_logits, end_logits, top_k, tau, gumbel, generator)
    idx = select_indices(cs, ce, of_lengths, pixel_values.shape[1], nframe, variant)
    return ops.gather_frames(pixel_values.contiguous(), idx), idx


# ------------------------------------------------------------------------------------------------ N2
def rouge_n(gold, pred, ignore=(",", ".")):
    """Host-side token-recall score of the self-refinement pass (src/gadgets/my_metrics.py:131-179), same return
    conventions: list inputs give one score per pair divided by len(gold) -- the reference's normalisation, kept
    as is -- and a list call with ignore=None returns after the first pair."""
    def recall(g, p):
        g, p = g.split(), p.split()
        if ignore is None:
            return sum(tok in p for tok in g) / len(g)
        kept = [tok for tok in g if tok not in ignore]
        return sum(tok in p for tok in kept) / len(kept) if kept else 0
    if type(gold) is list:
        out = []
        for g, p in zip(gold, pred):
            if ignore is None:
                return recall(g, p)
            r = recall(g, p)
            out.append(r / len(gold) if len(gold) > 0 else r)
        return out
    return recall(gold, pred)


def pseudo_span_targets(scores: torch.Tensor, of_lengths) -> tuple:
    """Per-frame scores [B, num_frames] (device, fp32) -> (start_targets, end_targets) int64 [B] on the device
    (LSTP_SF_blip2_module.py:195-224: monotone stack + rescale to flow positions; no `.tolist()` round trip)."""
    if not torch.is_tensor(of_lengths):
        of_lengths = torch.tensor(list(of_lengths), dtype=torch.int32, device=scores.device)
    return ops.pseudo_span(scores.to(torch.float32).contiguous(), of_lengths.to(torch.int32).contiguous())


def mrc_loss(of_logits: torch.Tensor, start_targets: torch.Tensor, end_targets: torch.Tensor) -> torch.Tensor:
    """Span cross-entropy on the sampler's mrc logits [B, Lof, 2] (LSTP_SF_blip2_module.py:238-253).
    FORWARD VALUE ONLY: the result carries no autograd graph (the kernels implement inference; training the sampler
    through this loss is out of scope) -- use it for evaluation / monitoring."""
    return ops.span_ce_loss(of_logits.to(torch.float32).contiguous(), start_targets.contiguous(), end_targets.contiguous())
