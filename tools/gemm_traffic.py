"""DRAM traffic of the ViT GEMM variants, one launch each (run under ncu):
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_bf16 \
        --csv --log-file gpurun_out/traffic.csv python tools/gemm_traffic.py
Prints the algorithmic bytes per launch for comparison."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from videotgb_b200 import ops

M = 257 * 512
SHAPES = [("proj", M, 1408, 1408, ops.EPI_F32_RESID), ("proj+stats", M, 1408, 1408, ops.EPI_F32_RESID_STATS),
          ("fc2", M, 1408, 6144, ops.EPI_F32_RESID), ("fc2+stats", M, 1408, 6144, ops.EPI_F32_RESID_STATS),
          ("qkv", M, 4224, 1408, ops.EPI_BF16), ("fc1", M, 6144, 1408, ops.EPI_BF16_GELU)]
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for name, m, n, k, epi in SHAPES:
    a = torch.randn(m, k, device="cuda").to(torch.bfloat16)
    w = (torch.randn(n, k, device="cuda") * 0.05).to(torch.bfloat16)
    bias = torch.randn(n, device="cuda")
    f32 = epi in (ops.EPI_F32_RESID, ops.EPI_F32_RESID_STATS)
    out = torch.randn(m, n, device="cuda") if f32 else torch.empty(m, n, dtype=torch.bfloat16, device="cuda")
    rd = (m * k + n * k) * 2 + (m * n * 4 if f32 else 0)
    wr = m * n * (4 if f32 else 2)
    flush.zero_()
    torch.cuda.synchronize()
    if epi == ops.EPI_F32_RESID_STATS:
        xb = torch.empty(m, n, dtype=torch.bfloat16, device="cuda")
        part = torch.empty(ops.ln_stats_slots(m, n), m, 2, device="cuda")
        wr += m * n * 2 + part.numel() * 4
        ops.gemm_ln(a, w, bias, epi, out, resid=out, out_bf16=xb, partial_stats=part)
    elif f32:
        ops.gemm(a, w, bias, epi, out=out, resid=out)
    else:
        ops.gemm(a, w, bias, epi, out=out)
    torch.cuda.synchronize()
    print(f"{name}: algorithmic read {rd / 1e9:.3f} GB write {wr / 1e9:.3f} GB", flush=True)
