"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the sampler's span draw, span->frame-index
selection and frame gather.

The reference has this logic inline (copy-pasted) in every task module; the files cannot be
imported here (Lightning etc. are absent), so this module restates it and
``oracle/gen_golden.py`` pins it by *executing the reference's own statements* read from
``/root/reference/src/models/LSTP_SF_blip2_module.py`` at generation time.

Variants (all from the reference):
  mapping "scaled":  int(s*(N-1)/(L-1))        LSTP_SF_blip2_module.py:446 (eval), :296 (train)
  mapping "ratio" :  int(s/L*N)                LSTP_module.py:245, eval/utils/model.py:135
  inclusive       :  range(start, end+1)       LSTP_SF_blip2_module.py:297 (train)
  exclusive       :  range(start, end)         LSTP_SF_blip2_module.py:447 (eval)

Numeric detail that matters for bit-exactness: when (s, e) come from ``argmax`` they are 0-dim int64
*tensors*, so ``s*(N-1)/(L-1)`` is a torch true-division evaluated in **float32**; when the fallback
replaces them with Python ints the same expression is evaluated in **float64**.  ``int()`` truncates.
"""
from __future__ import annotations

import numpy as np
import torch


def gumbel_argmax(start_logits: torch.Tensor, end_logits: torch.Tensor, gumbel: torch.Tensor, tau: float = 0.5):
    """One round of LSTP_SF_blip2_module.py:423-430 with the Gumbel noise made explicit.

    start/end logits [B,L,1] fp32, gumbel [2B,L,1] fp32 (= -log(Exp(1)) as F.gumbel_softmax draws it).
    Returns (start_idx [B], end_idx [B]) int64: argmax over dim=1 of softmax((logits+g)/tau).
    """
    logits = torch.cat([start_logits, end_logits], dim=0)
    y = ((logits + gumbel) / tau).softmax(dim=1)
    index = y.argmax(dim=1)                     # [2B,1]
    b = start_logits.shape[0]
    return index[:b, 0], index[b:, 0]


def _map_index(s: int, n: int, length: int, mapping: str, host_ints: bool) -> int:
    if mapping == "scaled":
        if host_ints:
            return int(s * (n - 1) / (length - 1))
        return int(np.float32(s * (n - 1)) / np.float32(length - 1))
    if mapping == "ratio":
        if host_ints:
            return int(s / length * n)
        return int(np.float32(np.float32(s) / np.float32(length)) * np.float32(n))
    raise ValueError(mapping)


def spans_to_indices(starts, ends, length: int, num_frames: int, nframe: int, mapping: str = "scaled",
                     inclusive: bool = False, host_ints: bool = False):
    """One video: the draws (starts[i], ends[i]) -> ``nframe`` sorted frame indices.

    Restates LSTP_SF_blip2_module.py:433-459.
    """
    cand = set()
    for s, e in zip(starts, ends):
        s, e = int(s), int(e)
        tensor_vals = not host_ints
        if s >= length or e >= length or (s == 0 and e == 0):
            s, e = 0, length - 1
            tensor_vals = False                              # Python ints from here on -> float64
        a = _map_index(s, num_frames, length, mapping, not tensor_vals)
        b = _map_index(e, num_frames, length, mapping, not tensor_vals)
        cand |= set(range(a, b + 1 if inclusive else b))
    cand = sorted(cand)
    if not cand:
        cand = list(range(num_frames))
    while len(cand) < nframe:
        cand = [xx for x in cand for xx in (x, x)]
    if len(cand) > nframe:
        intv = np.linspace(start=0, stop=len(cand), num=nframe + 1).astype(int)
        cand = [cand[(intv[x] + intv[x + 1] - 1) // 2] for x in range(len(intv) - 1)]
    assert len(cand) == nframe
    return cand


def uniform_indices(num_frames: int, nframe: int):
    """The uniform baseline (`cand_index = list(range(num_frames))`, LSTP_blip2_module.py:254)."""
    cand = list(range(num_frames))
    while len(cand) < nframe:
        cand = [xx for x in cand for xx in (x, x)]
    if len(cand) > nframe:
        intv = np.linspace(start=0, stop=len(cand), num=nframe + 1).astype(int)
        cand = [cand[(intv[x] + intv[x + 1] - 1) // 2] for x in range(len(intv) - 1)]
    return cand


def select_batch(cand_starts, cand_ends, of_lengths, num_frames: int, nframe: int, **kw):
    """cand_starts/ends: list over draws of [B] int tensors -> int64 [B, nframe]."""
    out = []
    for j, length in enumerate(of_lengths):
        out.append(spans_to_indices([int(c[j]) for c in cand_starts], [int(c[j]) for c in cand_ends],
                                    int(length), num_frames, nframe, **kw))
    return torch.tensor(out, dtype=torch.int64)


def gather_frames(pixel_values: torch.Tensor, indices: torch.Tensor) -> torch.Tensor:
    """[B,N,C,H,W] , [B,k] -> fp32 [B*k,C,H,W]   (LSTP_SF_blip2_module.py:417-419,460-464)."""
    b, n = pixel_values.shape[:2]
    k = indices.shape[1]
    out = torch.zeros((b, k) + tuple(pixel_values.shape[2:]), device=pixel_values.device)
    for j in range(b):
        out[j] = torch.index_select(pixel_values[j], 0, indices[j])
    return out.view((b * k,) + tuple(pixel_values.shape[2:]))


# --------------------------------------------------------------------------------------------
# self-refinement pseudo-labels (training forward)
#   src/models/LSTP_SF_blip2_module.py:195-227  per-frame scores -> monotone-stack span -> flow-index targets
#   src/models/LSTP_SF_blip2_module.py:241-253  span cross-entropy ("mrc") loss with ignore_index = Lof
# --------------------------------------------------------------------------------------------
def pseudo_span(scores, flow_length: int):
    """One video: per-frame scores (list of floats, as `tensor.tolist()` yields them) -> (start_target, end_target).

    Restates :197-217 (largest-rectangle monotone stack on the min-shifted, zero-padded scores) and the rescale of
    :221-222 (`int(t / (num_frames-1) * (flow_length-1))`, Python float64)."""
    num_frames = len(scores)
    bs = 0
    start_target, end_target = 0, num_frames - 1
    stack = []
    lo = min(scores)
    score = [0] + [s - lo for s in scores] + [0]
    for i in range(len(score)):
        while stack and score[stack[-1]] > score[i]:
            tmp = stack.pop()
            tmp_bs = (i - stack[-1] - 1) * score[tmp]
            if tmp_bs > bs:
                bs = tmp_bs
                start_target, end_target = stack[-1], i - 2
        stack.append(i)
    return (int(start_target / (num_frames - 1) * (flow_length - 1)),
            int(end_target / (num_frames - 1) * (flow_length - 1)))


def span_ce_loss(start_logits: torch.Tensor, end_logits: torch.Tensor, start_targets: torch.Tensor,
                 end_targets: torch.Tensor) -> torch.Tensor:
    """`mrc_loss` of :241-253: CrossEntropyLoss(ignore_index=Lof) on the squeezed start / end logits, averaged."""
    ignored_index = start_logits.size(1)
    loss_fct = torch.nn.CrossEntropyLoss(ignore_index=ignored_index)
    st = start_targets.clamp(0, ignored_index)
    et = end_targets.clamp(0, ignored_index)
    start_loss = loss_fct(start_logits.squeeze(-1).contiguous(), st)
    end_loss = loss_fct(end_logits.squeeze(-1).contiguous(), et)
    return (start_loss + end_loss) / 2
