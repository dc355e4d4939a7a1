"""Multi-GPU plumbing for the hot path.

The path shards by video / frame with no data-path collective until the very end: each rank scores, selects and
encodes its own videos, then the per-frame Q-Former tokens ([F_local, 32, 768], 48 KiB per frame in bf16) are
all-gathered over NVLink with NCCL so the rank that runs the unchanged reference LLM has every video's tokens.
The gather runs on a side stream and overlaps the local `language_projection` GEMM.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from . import ops


class TokenAllGather:
    """`exchange(query_output)` -> starts the all-gather of the bf16 tokens; `wait()` -> gathered [W*F_local, 32, 768]."""

    def __init__(self, group: Optional[dist.ProcessGroup] = None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.stream = torch.cuda.Stream() if torch.cuda.is_available() else None
        self.gathered = None
        self._work_done = None
        self.bytes_per_rank = 0

    def exchange(self, query_output: torch.Tensor) -> None:
        nf, nq, d = query_output.shape
        if query_output.is_cuda:
            local = ops.cast_bf16(query_output.to(torch.float32).contiguous().view(-1, d)).view(nf, nq, d)
        else:                                         # gloo path of the CPU unit tests: plumbing only
            local = query_output.to(torch.bfloat16).contiguous()
        self.bytes_per_rank = local.numel() * 2
        if self.gathered is None or self.gathered.shape[0] != self.world * nf:
            self.gathered = torch.empty((self.world * nf, nq, d), dtype=torch.bfloat16, device=local.device)
        if self.stream is not None and local.is_cuda:
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                dist.all_gather_into_tensor(self.gathered, local, group=self.group)
                self._work_done = torch.cuda.Event()
                self._work_done.record(self.stream)
            local.record_stream(self.stream)
        else:
            if local.dtype == torch.bfloat16 and not local.is_cuda:
                # gloo has no bf16 collectives: move the raw bytes
                g = self.gathered.view(torch.uint8)
                dist.all_gather_into_tensor(g, local.view(torch.uint8), group=self.group)
            else:
                dist.all_gather_into_tensor(self.gathered, local, group=self.group)

    def wait(self) -> torch.Tensor:
        if self._work_done is not None:
            torch.cuda.current_stream().wait_event(self._work_done)
            self._work_done = None
        return self.gathered


def shard_videos(num_videos: int, rank: int, world: int):
    """Contiguous block of videos for `rank` (videos are independent everywhere on the path)."""
    per = (num_videos + world - 1) // world
    lo = min(num_videos, rank * per)
    return lo, min(num_videos, lo + per)


# ------------------------------------------------------------------------------------------------ config 4
# One long video whose candidate frames / flow are sharded BY FRAME (contiguous ranges, rank r owns frames
# [r*T_local, (r+1)*T_local)): SURVEY.md section 8e.  Two exchange steps, both tiny next to the per-frame compute:
#   1. all-gather of the per-frame flow embeddings [B, T_local, 768] fp32 (3 KiB per frame) -- after it every rank runs
#      the small RoPE-BERT encoder replicated on identical inputs, so logits, the span draw (the Gumbel noise is
#      broadcast from rank 0) and the selected indices are identical everywhere with no further communication;
#   2. all-gather of the Q-Former tokens of the selected frames each rank owns, padded to `nframe` rows.
def all_gather_frame_rows(y_local: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """[B, T_local, D] per rank (same T_local everywhere) -> [B, W*T_local, D] in global frame order."""
    world = dist.get_world_size(group)
    if world == 1:
        return y_local
    b, tl, d = y_local.shape
    buf = torch.empty((world * b, tl, d), dtype=y_local.dtype, device=y_local.device)     # rank-major concatenation
    dist.all_gather_into_tensor(buf, y_local.contiguous(), group=group)
    return buf.view(world, b, tl, d).permute(1, 0, 2, 3).reshape(b, world * tl, d)


def owned_selection(indices: torch.Tensor, rank: int, t_local: int):
    """Sorted global frame indices [k] (replicated) -> (local indices this rank owns, position of the first one in the
    selection, per-rank counts list).  Shards are contiguous, so each rank owns a contiguous run of the sorted list."""
    idx = indices.reshape(-1).tolist()                     # one host sync per video (the reference syncs per comparison)
    owner = [i // t_local for i in idx]
    mine = [i - rank * t_local for i, o in zip(idx, owner) if o == rank]
    first = next((p for p, o in enumerate(owner) if o == rank), len(idx))
    world = max(owner) + 1 if owner else 1
    counts = [sum(1 for o in owner if o == r) for r in range(world)]
    return mine, first, counts


def all_gather_selected_tokens(tokens_local: torch.Tensor, counts, nframe: int,
                               group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """tokens_local [n_r, Q, D] (this rank's selected frames, in order; n_r = counts[rank]) -> [nframe, Q, D] on every
    rank, ordered by global frame index.  Every rank contributes a zero-padded [nframe, Q, D] slab."""
    world = dist.get_world_size(group)
    if world == 1:
        return tokens_local
    q, d = tokens_local.shape[1:]
    slab = torch.zeros((nframe, q, d), dtype=tokens_local.dtype, device=tokens_local.device)
    slab[: tokens_local.shape[0]] = tokens_local
    buf = torch.empty((world * nframe, q, d), dtype=slab.dtype, device=slab.device)
    dist.all_gather_into_tensor(buf, slab, group=group)
    counts = list(counts) + [0] * (world - len(counts))
    return torch.cat([buf[r * nframe: r * nframe + counts[r]] for r in range(world) if counts[r] > 0], dim=0)
