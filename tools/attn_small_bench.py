"""Q-Former attention shapes: cross and self on the tcgen05 kernels and on the generic mma.sync kernel: cross (32 queries x 257 image tokens) and self (32 x 32), 1024 frames.

    python tools/attn_small_bench.py       (B200 box; CUDA-graph replays, algorithmic bytes / time vs MEASURED_PEAKS hbm_gbs)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from videotgb_b200 import ops


def graph_time(fn, iters=10, replays=5):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (iters * replays)


def main():
    f, h, d = 1024, 12, 64
    q = torch.randn(f * 32, h * d, device="cuda").to(torch.bfloat16)
    kvs = [torch.randn(f * 257, 2 * h * d, device="cuda").to(torch.bfloat16) for _ in range(3)]     # > L2 together
    out = torch.empty(f * 32, h * d, dtype=torch.bfloat16, device="cuda")
    st = {"i": 0}

    def cross():
        kv = kvs[st["i"] % 3]
        st["i"] += 1
        ops.attention(q, kv[:, : h * d], kv[:, h * d:], f, h, 32, 257, d, d ** -0.5, out=out)

    nbytes = f * 257 * 2 * h * d * 2 + 2 * f * 32 * h * d * 2
    for tc in (True, False):
        ops.attention_set_tc(tc)
        ms = graph_time(cross)
        print(f"cross 32x257, {f} frames, {'tcgen05 ' if tc else 'mma.sync'}: {ms * 1e3:7.1f} us  {nbytes / ms / 1e6:7.0f} GB/s "
              f"({nbytes / ms / 1e6 / 6550.1:.2f} of the copy peak)")
    ops.attention_set_tc(True)
    if os.environ.get("TGB_QF_DBG"):       # timing experiments (wrong results): which resource bounds the tcgen05 kernel
        from videotgb_b200 import _lib
        lib = _lib.load()
        for bits, what in ((1, "no operand loads after the first ring fill"), (2, "no softmax arithmetic"),
                           (3, "neither loads nor softmax"), (4, "no output stores"), (0, "normal")):
            lib.tgb_debug_qf_attention_dbg(bits)
            ms = graph_time(cross)
            print(f"   dbg {bits} ({what}): {ms * 1e3:7.1f} us")
    qkv = torch.randn(f * 32, 3 * h * d, device="cuda").to(torch.bfloat16)
    nbytes = f * 32 * 4 * h * d * 2
    for tc in (True, False):
        ops.attention_set_tc(tc)
        ms = graph_time(lambda: ops.attention(qkv[:, : h * d], qkv[:, h * d: 2 * h * d], qkv[:, 2 * h * d:], f, h, 32, 32, d, d ** -0.5, out=out))
        print(f"self  32x32,  {f} frames, {'tcgen05 ' if tc else 'mma.sync'}: {ms * 1e3:7.1f} us  {nbytes / ms / 1e6:7.0f} GB/s "
              f"({nbytes / ms / 1e6 / 6550.1:.2f} of the copy peak; 201 MB, partly L2 resident)")
    ops.attention_set_tc(True)


if __name__ == "__main__":
    main()
