"""TEST INFRASTRUCTURE ONLY -- the seeded inputs of the golden model cases (shared by
``oracle/gen_golden.py`` and the tests so both sides see identical tensors)."""
import torch

from videotgb_b200 import synthetic as syn
from .transform_ref import normalize_u8_thwc


def vit_pixels(nf: int) -> torch.Tensor:
    u8 = syn.synthetic_frames_u8(nf, seed=11)
    return torch.from_numpy(normalize_u8_thwc(u8.numpy()))


def instruct_text(nf: int):
    g = torch.Generator().manual_seed(21)
    ids = torch.randint(1000, 30522, (1, 12), generator=g).repeat_interleave(nf, 0)
    tmask = torch.ones(nf, 12, dtype=torch.long)
    tmask[:, 9:] = 0
    amask = torch.cat([torch.ones(nf, 32, dtype=torch.long), tmask], dim=1)
    return ids, amask, g


def sampler_inputs():
    _, _, g = instruct_text(2)                      # same generator stream position as gen_golden
    b, lof, tq = 2, 12, 9
    of = syn.synthetic_flow(b, lof, seed=31)
    of_len = [12, 7]
    of_mask = torch.zeros(b, lof + 2, dtype=torch.long)
    for j, n in enumerate(of_len):
        of_mask[j, : n + 2] = 1
    qids = torch.randint(1000, 30522, (b, tq), generator=g)
    qmask = torch.ones(b, tq, dtype=torch.long)
    qmask[1, 6:] = 0
    return of, of_mask, of_len, qids, qmask


def raft_pair(n: int = 2, h: int = 128, w: int = 160, seed: int = 31):
    """A frame pair on the 0..255 scale: smooth random texture and a shifted, slightly noisy copy."""
    g = torch.Generator().manual_seed(seed)
    base = torch.nn.functional.interpolate(torch.rand(n, 3, h // 8, w // 8, generator=g), size=(h, w), mode="bicubic",
                                           align_corners=False).clamp(0, 1)
    img1 = base * 255.0
    img2 = (torch.roll(base, shifts=(2, -3), dims=(2, 3)) + 0.02 * torch.randn(n, 3, h, w, generator=g)).clamp(0, 1) * 255.0
    return img1.contiguous(), img2.contiguous()
