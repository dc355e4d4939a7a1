"""First-contact probe / micro-benchmark for the tcgen05 ViT attention kernel (run on the B200 box)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from videotgb_b200 import ops


def ref(qkv, f):
    x = qkv.float().view(f, 257, 3, 16, 88).permute(2, 0, 3, 1, 4)
    q, k, v = x[0], x[1], x[2]
    s = (q @ k.transpose(-1, -2)) * 88 ** -0.5
    return (s.softmax(-1) @ v).permute(0, 2, 1, 3).reshape(f * 257, 1408)


def probe(f, seed=0, scale_v=1.0):
    g = torch.Generator().manual_seed(seed)
    qkv = (torch.randn(f * 257, 4224, generator=g) * scale_v).to(torch.bfloat16).cuda()
    want = ref(qkv, f)
    try:
        got = ops.vit_attention(qkv, f, 88 ** -0.5).float()
        torch.cuda.synchronize()
    except Exception as e:                                   # noqa: BLE001
        print(f"[frames={f}] FAILED: {e}")
        return False
    err = (got - want).abs()
    rel = (err.max() / want.abs().max()).item()
    print(f"[frames={f}] max rel err {rel:.3e}  nan={torch.isnan(got).any().item()}")
    if rel > 2e-2 or torch.isnan(got).any():
        e3 = err.view(f, 257, 16, 88)
        print("   err by token group: cls", e3[:, 0].max().item(), " rows1-128", e3[:, 1:129].max().item(), " rows129-256", e3[:, 129:].max().item())
        print("   err by head:", [round(e3[:, :, h].max().item(), 3) for h in range(16)])
        print("   err by d block: d<48", e3[..., :48].max().item(), " 48<=d<64", e3[..., 48:64].max().item(), " d>=64", e3[..., 64:].max().item())
        print("   err by frame:", [round(e3[i].max().item(), 3) for i in range(min(f, 8))])
        print("   got[1, :8]", got.view(f, 257, 16, 88)[0, 1, 0, :8].tolist())
        print("   want[1,:8]", want.view(f, 257, 16, 88)[0, 1, 0, :8].tolist())
        # does it look like uniform attention (P constant)?  or unnormalised?
        ratio = (got.view(f, 257, 16, 88)[0, 1:, 0] / want.view(f, 257, 16, 88)[0, 1:, 0]).median().item()
        print("   median got/want ratio (head 0):", ratio)
        return False
    return True


def bench(f, iters=10):
    qkv = torch.randn(f * 257, 4224, device="cuda").to(torch.bfloat16)
    out = torch.empty(f * 257, 1408, dtype=torch.bfloat16, device="cuda")
    def variant(v):
        def run():
            ops.vit_attention_set_variant(v)
            ops.vit_attention(qkv, f, 88 ** -0.5, out=out)
            ops.vit_attention_set_variant(0)
        return run
    for name, fn in (("tcgen05 ping-pong + pipelined softmax", variant(5)),
                     ("tcgen05 split-row ping-pong", variant(4)),
                     ("tcgen05 CTA pair", variant(3)),
                     ("tcgen05 three-block", variant(2)),
                     ("tcgen05 ping-pong", variant(0)),
                     ("tcgen05 one-shot", variant(1)),
                     ("mma.sync", lambda: ops.attention(qkv[:, :1408], qkv[:, 1408:2816], qkv[:, 2816:], f, 16, 257, 257, 88, 88 ** -0.5, out=out))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        print(f"bench {name}: frames={f} {ms:.3f} ms  {4 * 257 * 257 * 88 * 16 * f / ms / 1e9:.0f} TFLOP/s (algorithmic)")


if __name__ == "__main__":
    ops.vit_attention_set_variant(int(os.environ.get("TGB_ATTN_VARIANT", "0")))
    ok = True
    for f in (1, 2, 3, 20):
        ok = probe(f, seed=f) and ok
    ok = probe(4, seed=9, scale_v=3.0) and ok
    if ok and "--bench" in sys.argv:
        bench(256)
        bench(512)
