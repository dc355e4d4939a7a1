"""ncu target: a few launches of the ViT attention kernel on 256 frames (no reference computation)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from videotgb_b200 import ops

f = int(sys.argv[1]) if len(sys.argv) > 1 else 256
qkv = torch.randn(f * 257, 4224, device="cuda").to(torch.bfloat16)
out = torch.empty(f * 257, 1408, dtype=torch.bfloat16, device="cuda")
for _ in range(4):
    ops.vit_attention(qkv, f, 88 ** -0.5, out=out)
torch.cuda.synchronize()
