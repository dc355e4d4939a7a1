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
