"""Event timeline of the ping-pong ViT attention kernel (CTA 0), to see which role waits on which.

    python tools/attn_trace.py build      (build container: compiles tools/_dbg/libtgb_trace.so with -DTGB_ATTN_TRACE)
    python tools/attn_trace.py [frames]   (B200 box: runs one launch, prints per-item events in SM clocks)
"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DBG = os.path.join(ROOT, "tools", "_dbg")
LIB = os.path.join(DBG, os.environ.get("TGB_TRACE_LIB", "libtgb_trace.so"))
CSRC = os.path.join(ROOT, "videotgb_b200", "csrc")

TAGS = {1: "TMA  K0 issued", 2: "TMA  Q0 issued", 3: "TMA  Q1 issued", 4: "TMA  V0 issued", 5: "TMA  K1 issued", 6: "TMA  V1 issued",
        10: "MMA  P(0,0) seen", 11: "MMA  P(1,0) seen", 12: "MMA  issue PV(0,0)+S(0,1)", 13: "MMA  issue PV(1,0)+S(1,1)",
        20: "MMA  P(0,1) seen -> PV(0,1)", 21: "MMA  P(1,1) seen -> PV(1,1)", 22: "MMA  next S(0,0) issue", 23: "MMA  next S(1,0) issue",
        30: "CLS  start", 31: "CLS  end",
        40: "WG   S(t,0) seen", 41: "WG   S(t,0) loaded", 42: "WG   exp done", 43: "WG   P(t,0) published",
        50: "WG   S(t,1) seen", 51: "WG   S(t,1) loaded", 52: "WG   exp done", 53: "WG   P(t,1) published",
        60: "WG   O seen", 61: "WG   O loaded", 62: "WG   stored"}
ROLE = ["tma", "mma", "cls2", "cls3", "wg0", "wg1"]

# three-block kernel (variant 2): unit r = 2*kb + t
P3_ROLE = ["tma", "mma", "epi", "-", "wg0", "wg1"]
P3_TAGS = {"tma": {1: "item loads start (K0)", 2: "Q0 issued", 3: "all loads issued"},
           "mma": {20 + r: f"PV(t={r & 1},kb={r >> 1}) + S(unit+3) issued" for r in range(6)},
           "epi": {60: "O_0 complete", 61: "O_1 complete", 62: "store 0 issued", 63: "store 1 issued"},
           "wg0": {}, "wg1": {}}
for _kb in range(3):
    for _wg in ("wg0", "wg1"):
        P3_TAGS[_wg].update({40 + 10 * _kb: f"S(t,{_kb}) seen", 42 + 10 * _kb: f"exp({_kb}) done", 43 + 10 * _kb: f"P(t,{_kb}) published"})

# CTA-pair kernel (variant 3), leader CTA of cluster 0
C2_ROLE = ["tma", "mma", "epi", "cls", "smA", "smB"]
C2_TAGS = {"tma": {1: "Q/K buffers free -> loads issued", 2: "V buffer free -> loads issued"},
           "mma": {10: "P(0) seen", 11: "O free", 12: "PV(0) issued", 13: "next item landed", 14: "next S(0) issued",
                   15: "P(1) seen", 16: "PV(1) issued", 17: "next S(1) issued"},
           "epi": {60: "O complete", 62: "store issued", 63: "CLS row done"},
           "cls": {30: "K landed", 31: "local scores sent", 32: "peer scores here", 33: "probabilities published"},
           "smA": {40: "S(0) seen", 41: "row max known", 42: "exp done", 43: "P(0) published", 50: "S(1) seen", 53: "P(1) published"},
           "smB": {40: "S(0) seen", 41: "row max known", 42: "exp done", 43: "P(0) published", 50: "S(1) seen", 53: "P(1) published"}}


def build():
    os.makedirs(DBG, exist_ok=True)
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]
    cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
           "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-DTGB_ATTN_TRACE", *os.environ.get("TGB_TRACE_DEFS", "").split(), "-shared", "-cudart", "static", "-o", LIB, *srcs]
    subprocess.run(cmd, check=True)
    print(LIB)


def run(frames, variant=0):
    import torch
    lib = ctypes.CDLL(LIB)
    lib.tgb_vit_attention_set_variant(variant)
    if os.environ.get("TGB_ATTN_PREFETCH"):
        lib.tgb_vit_attention_set_prefetch(int(os.environ["TGB_ATTN_PREFETCH"]))
    if os.environ.get("TGB_ATTN_DBG"):
        lib.tgb_debug_vit_attention_dbg(int(os.environ["TGB_ATTN_DBG"]))
    qkv = torch.randn(frames * 257, 4224, device="cuda").to(torch.bfloat16)
    out = torch.empty(frames * 257, 1408, dtype=torch.bfloat16, device="cuda")
    fn = lib.tgb_vit_attention_bf16
    fn.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_void_p]
    for _ in range(3):
        rc = fn(qkv.data_ptr(), 4224, out.data_ptr(), 1408, frames, 88 ** -0.5, None)
        assert rc == 0, rc
    torch.cuda.synchronize()
    cap = 1024
    buf = (ctypes.c_longlong * (6 * cap * 2))()
    n = (ctypes.c_int * 6)()
    assert lib.tgb_debug_attn_trace(buf, n) == cap
    ev = []
    for r in range(6):
        for k in range(n[r]):
            ev.append((buf[(r * cap + k) * 2 + 1], r, buf[(r * cap + k) * 2]))
    ev.sort()
    t0 = ev[0][0]
    last_item = max(e[2] // 100 for e in ev)
    for clk, r, tag in ev:
        item, code = tag // 100, tag % 100
        if item in (0, 1) or last_item // 2 <= item <= last_item // 2 + 1:
            if variant == 3:
                print(f"{clk - t0:9d}  item {item:3d}  {C2_ROLE[r]:5s} {C2_TAGS[C2_ROLE[r]].get(code, code)}")
            elif variant == 2:
                print(f"{clk - t0:9d}  item {item:3d}  {P3_ROLE[r]:5s} {P3_TAGS[P3_ROLE[r]].get(code, code)}")
            else:
                print(f"{clk - t0:9d}  item {item:3d}  {ROLE[r]:5s} {TAGS.get(code, code)}")
    # steady-state period
    ks = [clk for clk, r, tag in ev if r == 1 and tag % 100 == {2: 20, 3: 12}.get(variant, 12)]
    if len(ks) > 4:
        print("period (clk/item), items 2..:", (ks[-1] - ks[2]) / (len(ks) - 3))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "build":
        build()
    else:
        run(int(sys.argv[1]) if len(sys.argv) > 1 else 256, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
