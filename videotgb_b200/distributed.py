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
from .modeling import _cached_bf16


class TokenAllGather:
    """`exchange(query_output)` -> starts the all-gather of the bf16 tokens; `wait()` -> gathered [W*F_local, 32, 768]."""

    def __init__(self, group: Optional[dist.ProcessGroup] = None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.stream = torch.cuda.Stream() if torch.cuda.is_available() else None
        self.gathered = None
        self._work_done = None
        self.bytes_per_rank = 0

    def slot(self, nf: int, nq: int, d: int, device) -> torch.Tensor:
        """This rank's [nf, nq, d] bf16 slice of the gather buffer: the Q-Former's last LayerNorm writes the tokens here
        (`qformer(..., tokens_out_bf16=slot)`), so the collective starts from the kernel's own output -- no cast, no
        staging copy (in-place all-gather)."""
        if self.gathered is None or tuple(self.gathered.shape) != (self.world * nf, nq, d) or self.gathered.device != torch.device(device):
            self.gathered = torch.empty((self.world * nf, nq, d), dtype=torch.bfloat16, device=device)
        r = dist.get_rank(self.group)
        return self.gathered[r * nf: (r + 1) * nf]

    def exchange(self, query_output: torch.Tensor) -> None:
        nf, nq, d = query_output.shape
        pre = _cached_bf16(query_output)
        if query_output.dtype == torch.bfloat16:
            pre = query_output
        if pre is not None and self.gathered is not None and pre.data_ptr() == self.slot(nf, nq, d, pre.device).data_ptr():
            local = self.slot(nf, nq, d, pre.device)                 # already in place
        elif pre is not None:
            local = pre.contiguous()
        elif query_output.is_cuda:
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
#   2. all-to-all of the k selected u8 frames (150 KB each) so that every rank ENCODES k/W of them, whoever owns them --
#      selections are contiguous spans, so without this step one or two ranks would run the whole ViT;
#   3. all-gather of the bf16 Q-Former tokens of the frames each rank encoded (48 KiB per frame).
def all_gather_frame_rows(y_local: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """[B, T_local, D] per rank (same T_local everywhere) -> [B, W*T_local, D] in global frame order."""
    world = dist.get_world_size(group)
    if world == 1:
        return y_local
    b, tl, d = y_local.shape
    buf = torch.empty((world * b, tl, d), dtype=y_local.dtype, device=y_local.device)     # rank-major concatenation
    dist.all_gather_into_tensor(buf, y_local.contiguous(), group=group)
    return buf.view(world, b, tl, d).permute(1, 0, 2, 3).reshape(b, world * tl, d)


def encode_plan(indices, rank: int, world: int, t_local: int):
    """Who encodes which selected frame.  `indices`: the k sorted global frame indices of the selection (replicated on
    every rank, duplicates allowed).  The k selected POSITIONS are split evenly -- rank r encodes positions
    [r*k//W, (r+1)*k//W) whatever rank OWNS those candidates (selections are contiguous spans, so ownership is
    clustered on one or two ranks; SURVEY.md section 8e "contiguous blocks of the F = B*k selected frames per rank").
    Returns a namespace with
      lo, hi        the positions this rank encodes,
      send_local    local indices (into this rank's shard) of the selected frames it owns, in position order,
      send_splits   per destination rank: how many of those go there,
      recv_splits   per source rank: how many frames arrive from it (concatenated by source = position order),
      counts        per rank: hi - lo (for the token all-gather)."""
    from types import SimpleNamespace
    idx = [int(i) for i in indices]
    k = len(idx)
    bounds = [r * k // world for r in range(world + 1)]
    dest = [0] * k
    for r in range(world):
        for p in range(bounds[r], bounds[r + 1]):
            dest[p] = r
    owner = [i // t_local for i in idx]
    if owner and (min(owner) < 0 or max(owner) >= world):
        raise ValueError("selected frame index outside the sharded video")
    send_local = [idx[p] - rank * t_local for p in range(k) if owner[p] == rank]
    send_splits = [sum(1 for p in range(k) if owner[p] == rank and dest[p] == d) for d in range(world)]
    recv_splits = [sum(1 for p in range(k) if dest[p] == rank and owner[p] == s) for s in range(world)]
    return SimpleNamespace(lo=bounds[rank], hi=bounds[rank + 1], send_local=send_local, send_splits=send_splits,
                           recv_splits=recv_splits, counts=[bounds[r + 1] - bounds[r] for r in range(world)])


def redistribute_selected_frames(frames_local: torch.Tensor, plan, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """frames_local u8 [1, T_local, H, W, 3] (this rank's shard) -> u8 [hi-lo, H, W, 3]: the selected frames this rank
    ENCODES, in position order.  One variable-size all-to-all over NVLink (150 KB per frame); a rank's own frames move
    through the same call (device-local copy inside NCCL)."""
    shape = tuple(frames_local.shape[2:])
    n_send, n_recv = len(plan.send_local), plan.hi - plan.lo
    if n_send:
        local_idx = torch.tensor([plan.send_local], dtype=torch.int64, device=frames_local.device)
        send = ops.gather_frames(frames_local, local_idx) if frames_local.is_cuda else frames_local[0, local_idx[0]].contiguous()
    else:
        send = torch.empty((0,) + shape, dtype=frames_local.dtype, device=frames_local.device)
    recv = torch.empty((n_recv,) + shape, dtype=frames_local.dtype, device=frames_local.device)
    if dist.get_world_size(group) == 1:
        return send
    dist.all_to_all_single(recv, send, output_split_sizes=plan.recv_splits, input_split_sizes=plan.send_splits, group=group)
    return recv


def all_gather_encoded_tokens(tokens_local: torch.Tensor, counts, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """tokens_local [counts[rank], Q, D] -> [sum(counts), Q, D] on every rank in position order.  Equal counts (k % W == 0,
    the normal case) are one in-place all-gather with no padding and no re-assembly; otherwise slabs are padded to
    max(counts)."""
    world = dist.get_world_size(group)
    if world == 1:
        return tokens_local
    rank = dist.get_rank(group)
    q, d = tokens_local.shape[1:]
    cmax = max(counts)
    buf = torch.empty((world * cmax, q, d), dtype=tokens_local.dtype, device=tokens_local.device)
    slot = buf[rank * cmax: rank * cmax + counts[rank]]
    if slot.data_ptr() != tokens_local.data_ptr():            # the Q-Former may already have written into the slot
        slot.copy_(tokens_local)
    _all_gather_bytes(buf, buf[rank * cmax: (rank + 1) * cmax], group)
    if min(counts) == cmax:
        return buf
    return torch.cat([buf[r * cmax: r * cmax + counts[r]] for r in range(world) if counts[r] > 0], dim=0)


def _all_gather_bytes(out: torch.Tensor, local: torch.Tensor, group) -> None:
    if out.dtype == torch.bfloat16 and not out.is_cuda:        # gloo has no bf16 collectives: move the raw bytes
        dist.all_gather_into_tensor(out.view(torch.uint8), local.contiguous().view(torch.uint8), group=group)
    else:
        dist.all_gather_into_tensor(out, local, group=group)


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
