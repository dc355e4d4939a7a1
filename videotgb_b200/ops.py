"""Thin torch-tensor -> C-ABI call layer.  Each function validates dtype/device/contiguity, hands raw device
pointers plus the CURRENT torch CUDA stream to libvideotgb_b200.so, and returns torch tensors it allocated.
torch here is plumbing (device memory + streams); all arithmetic happens in the library's kernels.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _lib
from ._lib import check

EPI_BF16, EPI_BF16_GELU, EPI_F32, EPI_F32_RESID, EPI_F32_ROWBIAS, EPI_BF16_RELU = range(6)
EPI_BF16_LN, EPI_BF16_GELU_LN, EPI_F32_RESID_STATS = 6, 7, 8       # LayerNorm folded into the GEMMs (gemm_ln)
MAP_SCALED, MAP_RATIO = 0, 1
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


LAUNCHES = 0          # kernels launched through this layer (every C-ABI compute call enqueues exactly one kernel)
_PROFILE = None       # when a list: (name, shape, start_event, end_event) per launch, see profile_launches()


def _stream() -> int:
    global LAUNCHES
    LAUNCHES += 1
    return torch.cuda.current_stream().cuda_stream


class profile_launches:
    """Context manager: CUDA-event timing of every GEMM launch issued inside (used by bench.py's roofline leg)."""

    def __enter__(self):
        global _PROFILE
        _PROFILE = []
        return self

    def __exit__(self, *exc):
        global _PROFILE
        torch.cuda.synchronize()
        self.records = [(n, shp, a.elapsed_time(b)) for (n, shp, a, b) in _PROFILE]
        _PROFILE = None
        return False


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _need(t: torch.Tensor, dtype, name: str, contiguous: bool = True) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.TgbError(f"{name}: expected a CUDA tensor (the B200 path has no CPU fallback)")
    if t.dtype != dtype:
        raise _lib.TgbError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if t.device.index != torch.cuda.current_device():
        # kernels are enqueued on the CURRENT device's current stream: a tensor living elsewhere would be dereferenced
        # from the wrong device (wrap the call in `torch.cuda.device(t.device)`)
        raise _lib.TgbError(f"{name}: tensor is on cuda:{t.device.index} but the current device is "
                            f"cuda:{torch.cuda.current_device()}")
    if contiguous and not t.is_contiguous():
        raise _lib.TgbError(f"{name}: expected a contiguous tensor")
    return t


def _f3(vals):
    return (ctypes.c_float * 3)(*[float(v) for v in vals])


def _rows(t: torch.Tensor, name: str):
    """2-D row-major view parameters (rows, cols, leading dim) of a tensor whose last dim is contiguous."""
    if t.dim() < 2 or t.stride(-1) != 1:
        raise _lib.TgbError(f"{name}: last dimension must be contiguous")
    if t.dim() == 2:
        return t.shape[0], t.shape[1], t.stride(0)
    if not t.is_contiguous():
        raise _lib.TgbError(f"{name}: >2-D tensors must be contiguous")
    return t.numel() // t.shape[-1], t.shape[-1], t.shape[-1]


# ------------------------------------------------------------------------------------------------ A0
def normalize_frames_u8(frames_thwc: torch.Tensor, mean=CLIP_MEAN, std=CLIP_STD) -> torch.Tensor:
    f = _need(frames_thwc, torch.uint8, "frames")
    t, h, w, c = f.shape
    if c != 3:
        raise _lib.TgbError("frames must be [T,H,W,3]")
    out = torch.empty((t, 3, h, w), dtype=torch.float32, device=f.device)
    check(_lib.load().tgb_normalize_frames_u8(_p(f), _p(out), t, h, w, _f3(mean), _f3(std), _stream()))
    return out


def resize_bilinear_u8(frames_thwc: torch.Tensor, size: int) -> torch.Tensor:
    """u8 [T,H0,W0,3] decoded frames -> u8 [T,size,size,3] (bilinear, align_corners=False, truncation)."""
    f = _need(frames_thwc, torch.uint8, "frames")
    if f.dim() != 4 or f.shape[3] != 3:
        raise _lib.TgbError("frames must be [T,H0,W0,3]")
    t, h0, w0, _ = f.shape
    out = torch.empty((t, size, size, 3), dtype=torch.uint8, device=f.device)
    for lo in range(0, t, 65535):
        n = min(65535, t - lo)
        check(_lib.load().tgb_resize_bilinear_u8(_p(f[lo:]), _p(out[lo:]), n, h0, w0, int(size), _stream()))
    return out


def resize_set_variant(variant: int) -> None:
    """0 = auto (word-load kernel for 4-byte aligned inputs), 1 = generic byte-load kernel; test / debug knob."""
    check(_lib.load().tgb_resize_set_variant(int(variant)))


# ------------------------------------------------------------------------------------------------ A7-A9
def gumbel_argmax(logits: torch.Tensor, gumbel: torch.Tensor, tau: float = 0.5) -> torch.Tensor:
    lg = _need(logits, torch.float32, "logits")
    gm = _need(gumbel, torch.float32, "gumbel")
    if lg.shape != gm.shape or lg.dim() != 2:
        raise _lib.TgbError("logits and gumbel must both be [R, L]")
    idx = torch.empty((lg.shape[0],), dtype=torch.int64, device=lg.device)
    check(_lib.load().tgb_gumbel_argmax(_p(lg), _p(gm), float(tau), lg.shape[0], lg.shape[1], _p(idx), _stream()))
    return idx


def additive_mask(mask: torch.Tensor, value: float, want_out: bool = True, want_last_index: bool = False):
    """0/1 mask [B, S] -> (fp32 additive mask (1 - mask) * value or None, int32 sum(mask) - 1 per row or None), one launch."""
    if not isinstance(mask, torch.Tensor) or not mask.is_cuda or mask.dim() != 2:
        raise _lib.TgbError("mask: expected a CUDA tensor [B, S]")
    m = mask if mask.dtype == torch.int64 else mask.to(torch.int64)
    m = _need(m.contiguous(), torch.int64, "mask")
    b, s = m.shape
    out = torch.empty((b, s), dtype=torch.float32, device=m.device) if want_out else None
    last = torch.empty((b,), dtype=torch.int32, device=m.device) if want_last_index else None
    check(_lib.load().tgb_additive_mask(_p(m), b, s, float(value), _p(out), _p(last), _stream()))
    return out, last


def span_draw(start_logits: torch.Tensor, end_logits: torch.Tensor, noise: torch.Tensor, noise_is_exponential: bool,
              tau: float, cand_start: torch.Tensor, cand_end: torch.Tensor) -> None:
    """One draw round: start/end logits f32 [B, L], noise f32 [2B, L] -> cand_start / cand_end i64 [B] (written in place,
    typically rows of the [top_k, B] candidate lists)."""
    sl = _need(start_logits, torch.float32, "start_logits")
    el = _need(end_logits, torch.float32, "end_logits")
    nz = _need(noise, torch.float32, "noise")
    b, l = sl.shape
    if el.shape != sl.shape or nz.shape != (2 * b, l):
        raise _lib.TgbError("start/end logits must be [B, L] and the noise [2B, L]")
    cs = _need(cand_start, torch.int64, "cand_start")
    ce = _need(cand_end, torch.int64, "cand_end")
    if cs.shape != (b,) or ce.shape != (b,):
        raise _lib.TgbError("cand_start / cand_end must be [B]")
    check(_lib.load().tgb_span_draw(_p(sl), _p(el), _p(nz), int(bool(noise_is_exponential)), float(tau), b, l, _p(cs), _p(ce),
                                    _stream()))


def spans_to_indices(cand_start: torch.Tensor, cand_end: torch.Tensor, of_lengths: torch.Tensor, num_frames: int,
                     nframe: int, mapping: int = MAP_SCALED, inclusive: bool = False,
                     host_ints: bool = False) -> torch.Tensor:
    cs = _need(cand_start, torch.int64, "cand_start")
    ce = _need(cand_end, torch.int64, "cand_end")
    ol = _need(of_lengths, torch.int32, "of_lengths")
    if cs.dim() != 2 or cs.shape != ce.shape or ol.shape != (cs.shape[1],):
        raise _lib.TgbError("cand_start/cand_end must be [top_k, B] and of_lengths [B]")
    out = torch.empty((cs.shape[1], nframe), dtype=torch.int64, device=cs.device)
    check(_lib.load().tgb_spans_to_indices(_p(cs), _p(ce), cs.shape[0], _p(ol), cs.shape[1], int(num_frames),
                                           int(nframe), int(mapping), int(bool(inclusive)), int(bool(host_ints)),
                                           _p(out), _stream()))
    return out


def pseudo_span(scores: torch.Tensor, of_lengths: torch.Tensor):
    """f32 [B, num_frames] per-frame scores, i32 [B] flow lengths -> (start_targets, end_targets) i64 [B]."""
    sc = _need(scores, torch.float32, "scores")
    ol = _need(of_lengths, torch.int32, "of_lengths")
    if sc.dim() != 2 or ol.shape != (sc.shape[0],):
        raise _lib.TgbError("scores must be [B, num_frames] and of_lengths [B]")
    st = torch.empty((sc.shape[0],), dtype=torch.int64, device=sc.device)
    en = torch.empty_like(st)
    check(_lib.load().tgb_pseudo_span(_p(sc), _p(ol), sc.shape[0], sc.shape[1], _p(st), _p(en), _stream()))
    return st, en


def span_ce_loss(logits: torch.Tensor, start_targets: torch.Tensor, end_targets: torch.Tensor) -> torch.Tensor:
    """mrc logits f32 [B, L, 2], i64 [B] targets -> 0-dim f32 loss."""
    lg = _need(logits, torch.float32, "logits")
    st = _need(start_targets, torch.int64, "start_targets")
    en = _need(end_targets, torch.int64, "end_targets")
    if lg.dim() != 3 or lg.shape[2] != 2 or st.shape != (lg.shape[0],) or en.shape != st.shape:
        raise _lib.TgbError("logits must be [B, L, 2] and the targets [B]")
    out = torch.empty((1,), dtype=torch.float32, device=lg.device)
    check(_lib.load().tgb_span_ce_loss(_p(lg), _p(st), _p(en), lg.shape[0], lg.shape[1], _p(out), _stream()))
    return out[0]


def gather_frames(pixel_values: torch.Tensor, indices: torch.Tensor) -> torch.Tensor:
    """[B,N,...] , i64 [B,k] -> [B*k, ...] (same dtype as the source; the reference buffer is fp32)."""
    if not pixel_values.is_cuda or not pixel_values.is_contiguous():
        raise _lib.TgbError("pixel_values must be a contiguous CUDA tensor")
    idx = _need(indices, torch.int64, "indices")
    b, n = pixel_values.shape[:2]
    k = idx.shape[1]
    frame_bytes = pixel_values[0, 0].numel() * pixel_values.element_size()
    out = torch.empty((b * k,) + tuple(pixel_values.shape[2:]), dtype=pixel_values.dtype, device=pixel_values.device)
    check(_lib.load().tgb_gather_frames(_p(pixel_values), _p(idx), _p(out), b, n, k, frame_bytes, _stream()))
    return out


# ------------------------------------------------------------------------------------------------ A10
def patchify(pixel_values: torch.Tensor, patch: int, ld: int) -> torch.Tensor:
    px = _need(pixel_values, torch.float32, "pixel_values")
    f, c, s, s2 = px.shape
    if c != 3 or s != s2:
        raise _lib.TgbError("pixel_values must be [F,3,S,S]")
    out = torch.empty((f * (s // patch) ** 2, ld), dtype=torch.bfloat16, device=px.device)
    check(_lib.load().tgb_patchify_f32(_p(px), _p(out), f, s, patch, ld, _stream()))
    return out


def patchify_u8(frames_thwc: torch.Tensor, patch: int, ld: int, mean=CLIP_MEAN, std=CLIP_STD) -> torch.Tensor:
    fr = _need(frames_thwc, torch.uint8, "frames")
    f, s, s2, c = fr.shape
    if c != 3 or s != s2:
        raise _lib.TgbError("frames must be [F,S,S,3]")
    out = torch.empty((f * (s // patch) ** 2, ld), dtype=torch.bfloat16, device=fr.device)
    check(_lib.load().tgb_patchify_u8(_p(fr), _p(out), f, s, patch, ld, _f3(mean), _f3(std), _stream()))
    return out


def patchify_set_variant(variant: int) -> None:
    """Test knob: 1 keeps u8 frames off the single-fma pair kernel (table / residual-corrected kernels instead)."""
    check(_lib.load().tgb_patchify_set_variant(int(variant)))


def fill_cls_rows(hidden: torch.Tensor, cls: torch.Tensor, pos0: torch.Tensor) -> None:
    h = _need(hidden, torch.float32, "hidden")
    f, tokens, d = h.shape
    check(_lib.load().tgb_fill_cls_rows(_p(h), _p(_need(cls, torch.float32, "cls")), _p(_need(pos0, torch.float32, "pos0")),
                                        f, tokens, d, _stream()))


# ------------------------------------------------------------------------------------------------ GEMM
def gemm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, epilogue: int = EPI_BF16,
         out: Optional[torch.Tensor] = None, resid: Optional[torch.Tensor] = None,
         rowbias: Optional[torch.Tensor] = None, period: int = 0, skip: int = 0, bn: int = 0) -> torch.Tensor:
    """out = epilogue(a @ w.T).  a bf16 [M,K] (row stride allowed), w bf16 [N,K]."""
    _need(a, torch.bfloat16, "a", contiguous=False)
    _need(w, torch.bfloat16, "w", contiguous=False)
    m, k, lda = _rows(a, "a")
    n, kw, ldw = _rows(w, "w")
    if k != kw:
        raise _lib.TgbError(f"gemm: K mismatch {k} vs {kw}")
    out_dtype = torch.bfloat16 if epilogue in (EPI_BF16, EPI_BF16_GELU, EPI_BF16_RELU) else torch.float32
    if out is None:
        if epilogue == EPI_F32_ROWBIAS:
            raise _lib.TgbError("gemm: the row-bias epilogue writes into a caller-provided buffer")
        out = torch.empty((m, n), dtype=out_dtype, device=a.device)
    _need(out, out_dtype, "out", contiguous=False)
    _, n_out, ldo = _rows(out, "out")
    if n_out != n:
        raise _lib.TgbError("gemm: out has the wrong width")
    ldr = 0
    if resid is not None:
        _need(resid, torch.float32, "resid", contiguous=False)
        _, _, ldr = _rows(resid, "resid")
    if bias is not None:
        _need(bias, torch.float32, "bias")
    if rowbias is not None:
        _need(rowbias, torch.float32, "rowbias")
    if _PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(_lib.load().tgb_gemm_bf16(_p(a), lda, _p(w), ldw, m, n, k, epilogue, _p(bias), _p(resid), ldr, _p(rowbias),
                                    period, skip, _p(out), ldo, bn, _stream()))
    if _PROFILE is not None:
        e1.record()
        _PROFILE.append(("gemm", (m, n, k, epilogue), e0, e1))
    return out


def gemm_ln(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, epilogue: int, out: torch.Tensor,
            rowstats: Optional[torch.Tensor] = None, colsum: Optional[torch.Tensor] = None,
            resid: Optional[torch.Tensor] = None, out_bf16: Optional[torch.Tensor] = None,
            partial_stats: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The LayerNorm-fold GEMMs (include/videotgb_b200.h, tgb_gemm_bf16_ln).
    EPI_BF16_LN / EPI_BF16_GELU_LN: out bf16 = act(rstd_m * (a @ w.T - mean_m * colsum) + bias), rowstats f32 [M,2].
    EPI_F32_RESID_STATS: out f32 = a @ w.T + bias + resid, out_bf16 = bf16(out), partial_stats f32 [slots, M, 2] (slot-major)."""
    _need(a, torch.bfloat16, "a", contiguous=False)
    _need(w, torch.bfloat16, "w", contiguous=False)
    m, k, lda = _rows(a, "a")
    n, kw, ldw = _rows(w, "w")
    if k != kw:
        raise _lib.TgbError(f"gemm_ln: K mismatch {k} vs {kw}")
    producer = epilogue == EPI_F32_RESID_STATS
    _need(out, torch.float32 if producer else torch.bfloat16, "out", contiguous=False)
    mo, n_out, ldo = _rows(out, "out")
    if n_out != n or mo != m:
        raise _lib.TgbError("gemm_ln: out has the wrong shape")
    _need(bias, torch.float32, "bias")
    ldr = ldo2 = 0
    if producer:
        _need(resid, torch.float32, "resid", contiguous=False)
        _, _, ldr = _rows(resid, "resid")
        _need(out_bf16, torch.bfloat16, "out_bf16", contiguous=False)
        m2, n2, ldo2 = _rows(out_bf16, "out_bf16")
        _need(partial_stats, torch.float32, "partial_stats")
        if (m2, n2) != (m, n) or partial_stats.numel() != m * ln_stats_slots(m, n) * 2:
            raise _lib.TgbError("gemm_ln: out_bf16 / partial_stats have the wrong shape")
    else:
        _need(rowstats, torch.float32, "rowstats")
        _need(colsum, torch.float32, "colsum")
        if rowstats.numel() != 2 * m or colsum.numel() != n:
            raise _lib.TgbError("gemm_ln: rowstats must be [M,2] and colsum [N]")
    if _PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(_lib.load().tgb_gemm_bf16_ln(_p(a), lda, _p(w), ldw, m, n, k, epilogue, _p(bias), _p(resid), ldr, _p(rowstats),
                                       _p(colsum), _p(out), ldo, _p(out_bf16), ldo2, _p(partial_stats), _stream()))
    if _PROFILE is not None:
        e1.record()
        _PROFILE.append(("gemm", (m, n, k, epilogue), e0, e1))
    return out


def ln_stats_slots(m: int, n: int) -> int:
    """Partial-statistics slots per row the EPI_F32_RESID_STATS epilogue writes for an [m, n] output."""
    return int(_lib.load().tgb_gemm_ln_stats_slots(int(m), int(n)))


def ln_stats_finalize(partial_stats: torch.Tensor, n: int, eps: float, rowstats: torch.Tensor,
                      max_ratio: Optional[torch.Tensor] = None) -> torch.Tensor:
    """partial_stats f32 [slots, M, 2] (sum, sum of squares; slot-major) -> rowstats f32 [M, 2] (mean, rstd).
    max_ratio: optional f32 scalar tensor that accumulates max |mean| * rstd over the rows (zero it first)."""
    _need(partial_stats, torch.float32, "partial_stats")
    _need(rowstats, torch.float32, "rowstats")
    slots, m, two = partial_stats.shape
    if two != 2 or rowstats.shape != (m, 2):
        raise _lib.TgbError("ln_stats_finalize: partial_stats must be [slots, M, 2] and rowstats [M, 2]")
    if _PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    if max_ratio is not None:
        _need(max_ratio, torch.float32, "max_ratio")
    check(_lib.load().tgb_ln_stats_finalize(_p(partial_stats), m, slots, int(n), float(eps), _p(rowstats), _p(max_ratio),
                                            _stream()))
    if _PROFILE is not None:
        e1.record()
        _PROFILE.append(("layernorm", (m, n, "finalize"), e0, e1))
    return rowstats


def rowstats_cast(x: torch.Tensor, eps: float, out_bf16: torch.Tensor, rowstats: torch.Tensor) -> None:
    """x f32 [M,N] -> bf16 copy + (mean, rstd) per row (LayerNorm fold, rows that did not come out of a GEMM)."""
    _need(x, torch.float32, "x", contiguous=False)
    m, n, ldx = _rows(x, "x")
    _need(out_bf16, torch.bfloat16, "out_bf16", contiguous=False)
    m2, n2, ldo = _rows(out_bf16, "out_bf16")
    _need(rowstats, torch.float32, "rowstats")
    if (m2, n2) != (m, n) or rowstats.shape != (m, 2):
        raise _lib.TgbError("rowstats_cast: shape mismatch")
    if _PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(_lib.load().tgb_rowstats_cast_f32(_p(x), ldx, m, n, float(eps), _p(out_bf16), ldo, _p(rowstats), _stream()))
    if _PROFILE is not None:
        e1.record()
        _PROFILE.append(("layernorm", (m, n, "rowstats"), e0, e1))


def gemm_set_cta_group(cta_group: int) -> None:
    """0 = auto, 1 = single-CTA tiles, 2 = CTA-pair (cta_group::2) tiles; test / debug knob."""
    check(_lib.load().tgb_gemm_set_cta_group(int(cta_group)))


def gemm_set_schedule(mode) -> None:
    """1 / True = balanced persistent tile schedule (default); 0 / False = plain round-robin; 2 = whole A row blocks per
    cluster when M allows (measured slower); 3 = fixed groups of N-block-count clusters sharing an A row block.  A/B knob."""
    check(_lib.load().tgb_gemm_set_schedule(1 if mode is True else int(mode)))


def gemm_set_epilogue_warps(warps: int) -> None:
    """0 = auto (8), 8 or 16 (the fp32-residual epilogues always run 8 warps); test / debug knob."""
    check(_lib.load().tgb_gemm_set_epilogue_warps(int(warps)))


def gemm_set_resid_piece(width: int) -> None:
    """0 = auto (32-column pieces when K <= 2048, else 16), 16 or 32; test / A-B knob of the fp32-residual epilogues."""
    check(_lib.load().tgb_gemm_set_resid_piece(int(width)))


# ------------------------------------------------------------------------------------------------ LayerNorm
def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, want_bf16: bool = True,
              want_f32: bool = False, out_bf16: Optional[torch.Tensor] = None, out_f32: Optional[torch.Tensor] = None):
    _need(x, torch.float32, "x", contiguous=False)
    m, n, ldx = _rows(x, "x")
    ob = out_bf16 if out_bf16 is not None else (
        torch.empty((m, n), dtype=torch.bfloat16, device=x.device) if want_bf16 else None)
    of = out_f32 if out_f32 is not None else (
        torch.empty((m, n), dtype=torch.float32, device=x.device) if want_f32 else None)
    if ob is not None and (ob.dtype != torch.bfloat16 or ob.shape != (m, n) or not ob.is_contiguous()):
        raise _lib.TgbError("layernorm: out_bf16 must be a contiguous bf16 [M,N] tensor")
    if of is not None and (of.dtype != torch.float32 or of.shape != (m, n) or not of.is_contiguous()):
        raise _lib.TgbError("layernorm: out_f32 must be a contiguous f32 [M,N] tensor")
    if _PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(_lib.load().tgb_layernorm_f32(_p(x), ldx, _p(_need(gamma, torch.float32, "gamma")),
                                        _p(_need(beta, torch.float32, "beta")), float(eps), m, n, _p(ob), n, _p(of), n,
                                        _stream()))
    if _PROFILE is not None:
        e1.record()
        _PROFILE.append(("layernorm", (m, n, ob is not None, of is not None), e0, e1))
    return ob, of


# ------------------------------------------------------------------------------------------------ attention
def self_attention_2seg(qkv1: torch.Tensor, qkv2: torch.Tensor, batch: int, heads: int, n1: int, n2: int, head_dim: int,
                        scale: float, mask: Optional[torch.Tensor] = None):
    """Self-attention over cat(rows of qkv1, rows of qkv2) per batch element without materialising the concatenation.
    qkv1 bf16 [batch*n1, 3*heads*head_dim], qkv2 [batch*n2, ...] -> (out1 [batch*n1, heads*head_dim], out2 [batch*n2, ...])."""
    hd = heads * head_dim
    for t, nm, n in ((qkv1, "qkv1", n1), (qkv2, "qkv2", n2)):
        _need(t, torch.bfloat16, nm)
        if t.dim() != 2 or t.shape != (batch * n, 3 * hd):
            raise _lib.TgbError(f"self_attention_2seg: {nm} must be [batch*n, 3*heads*head_dim]")
    if mask is not None:
        _need(mask, torch.float32, "mask")
        if mask.shape != (batch, n1 + n2):
            raise _lib.TgbError("self_attention_2seg: mask must be [batch, n1+n2]")
    out1 = torch.empty((batch * n1, hd), dtype=torch.bfloat16, device=qkv1.device)
    out2 = torch.empty((batch * n2, hd), dtype=torch.bfloat16, device=qkv1.device)
    check(_lib.load().tgb_self_attention_2seg_bf16(_p(qkv1), _p(qkv2), 3 * hd, _p(out1), _p(out2), hd, _p(mask), batch, heads,
                                                   n1, n2, head_dim, float(scale), _stream()))
    return out1, out2


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, batch: int, heads: int, sq: int, sk: int,
              head_dim: int, scale: float, mask: Optional[torch.Tensor] = None,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """q/k/v: bf16 2-D views [batch*S, >= heads*head_dim] (column slices of a fused projection are fine)."""
    for t, nm in ((q, "q"), (k, "k"), (v, "v")):
        _need(t, torch.bfloat16, nm, contiguous=False)
        if t.dim() != 2 or t.stride(1) != 1:
            raise _lib.TgbError(f"attention: {nm} must be a 2-D row-major view")
    if q.shape[0] != batch * sq or k.shape[0] != batch * sk or v.shape[0] != batch * sk:
        raise _lib.TgbError("attention: row counts do not match batch*S")
    if out is None:
        out = torch.empty((batch * sq, heads * head_dim), dtype=torch.bfloat16, device=q.device)
    if mask is not None:
        _need(mask, torch.float32, "mask")
        if mask.shape != (batch, sk):
            raise _lib.TgbError("attention: mask must be [batch, Sk]")
    if _PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(_lib.load().tgb_attention_bf16(_p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(out),
                                         out.stride(0), _p(mask), batch, heads, sq, sk, head_dim, float(scale), _stream()))
    if _PROFILE is not None:
        e1.record()
        _PROFILE.append(("attention", (batch, heads, sq, sk, head_dim), e0, e1))
    return out


# ------------------------------------------------------------------------------------------------ misc
def cast_bf16(x: torch.Tensor) -> torch.Tensor:
    _need(x, torch.float32, "x", contiguous=False)
    m, n, ldx = _rows(x, "x")
    out = torch.empty((m, n), dtype=torch.bfloat16, device=x.device)
    check(_lib.load().tgb_cast_f32_bf16(_p(x), ldx, _p(out), n, m, n, _stream()))
    return out


def group_mean(x: torch.Tensor, groups: int, width: int) -> torch.Tensor:
    _need(x, torch.float32, "x")
    if x.shape[0] != groups * width:
        raise _lib.TgbError("group_mean: leading dim must be groups*width")
    inner = x[0].numel()
    out = torch.empty((groups,) + tuple(x.shape[1:]), dtype=torch.float32, device=x.device)
    check(_lib.load().tgb_group_mean_f32(_p(x), _p(out), groups, width, inner, _stream()))
    return out


# ------------------------------------------------------------------------------------------------ sampler
def of_patch_reduce(flow: torch.Tensor, fc_weight: torch.Tensor, patch: int = 16) -> torch.Tensor:
    """flow f32 [T, C, S, S] -> bf16 [T, C*patch*patch] weighted patch reduction (see csrc/sampler.cu)."""
    fl = _need(flow, torch.float32, "flow")
    t, c, s, s2 = fl.shape
    fw = _need(fc_weight, torch.float32, "fc_weight")
    if s != s2 or fw.numel() != (s // patch) ** 2:
        raise _lib.TgbError("of_patch_reduce: fc_weight must have (S/patch)^2 elements")
    z = torch.empty((t, c * patch * patch), dtype=torch.bfloat16, device=fl.device)
    check(_lib.load().tgb_of_patch_reduce(_p(fl), _p(fw), _p(z), t, c, s, patch, z.shape[1], _stream()))
    return z


def sampler_assemble(y: torch.Tensor, bos, eos, pos, ends: torch.Tensor, gamma, beta, eps: float):
    """y f32 [B, L, N] -> (f32, bf16) [B, L+2, N]."""
    _need(y, torch.float32, "y")
    b, l, n = y.shape
    _need(ends, torch.int32, "ends")
    of = torch.empty((b, l + 2, n), dtype=torch.float32, device=y.device)
    ob = torch.empty((b, l + 2, n), dtype=torch.bfloat16, device=y.device)
    if pos.shape[0] < l + 2:
        raise _lib.TgbError(f"sampler: {l + 2} positions exceed max_position_embeddings={pos.shape[0]}")
    check(_lib.load().tgb_sampler_assemble(_p(y), _p(_need(bos, torch.float32, "bos")), _p(_need(eos, torch.float32, "eos")),
                                           _p(_need(pos, torch.float32, "pos")), _p(ends), _p(_need(gamma, torch.float32, "gamma")),
                                           _p(_need(beta, torch.float32, "beta")), float(eps), b, l, n, _p(of), _p(ob), _stream()))
    return of, ob


def embed_layernorm(ids: torch.Tensor, table: torch.Tensor, gamma, beta, eps: float, pos: Optional[torch.Tensor] = None,
                    period: int = 0, addvec: Optional[torch.Tensor] = None, out_f32: Optional[torch.Tensor] = None,
                    out_bf16: Optional[torch.Tensor] = None):
    """ids i64 [R] -> LayerNorm(table[ids] + pos[r % period] + addvec) written to 2-D (possibly strided) outputs."""
    _need(ids, torch.int64, "ids")
    _need(table, torch.float32, "table")
    r, n = ids.numel(), table.shape[1]
    if out_f32 is None and out_bf16 is None:
        out_f32 = torch.empty((r, n), dtype=torch.float32, device=ids.device)
        out_bf16 = torch.empty((r, n), dtype=torch.bfloat16, device=ids.device)
    ldf = out_f32.stride(0) if out_f32 is not None else 0
    ldb = out_bf16.stride(0) if out_bf16 is not None else 0
    check(_lib.load().tgb_embed_layernorm(_p(ids), _p(table), table.shape[0], _p(pos), int(period), _p(addvec),
                                          _p(_need(gamma, torch.float32, "gamma")), _p(_need(beta, torch.float32, "beta")),
                                          float(eps), r, n, _p(out_f32), ldf, _p(out_bf16), ldb, _stream()))
    return out_f32, out_bf16


def rope_(x: torch.Tensor, table: torch.Tensor, seq: int, heads: int, head_dim: int) -> torch.Tensor:
    """In-place rotary embedding on a bf16 2-D view [rows, >= heads*head_dim]; position = row % seq."""
    _need(x, torch.bfloat16, "x", contiguous=False)
    _need(table, torch.float32, "table")
    if x.dim() != 2 or x.stride(1) != 1 or table.shape[0] < seq or table.shape[1] != head_dim:
        raise _lib.TgbError("rope: bad shapes")
    check(_lib.load().tgb_rope_bf16(_p(x), x.stride(0), _p(table), x.shape[0], seq, heads, head_dim, _stream()))
    return x


def linear_small(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """fp32 out[M,N] = x[M,K] @ w[N,K].T + bias for tiny N (<= 8)."""
    _need(x, torch.float32, "x", contiguous=False)
    m, k, ldx = _rows(x, "x")
    _need(w, torch.float32, "w")
    n = w.shape[0]
    out = torch.empty((m, n), dtype=torch.float32, device=x.device)
    check(_lib.load().tgb_linear_small_f32(_p(x), ldx, _p(w), _p(bias), m, n, k, _p(out), n, _stream()))
    return out


def attention_set_tc(on: bool) -> None:
    """Test knob: True (default) = the Q-Former cross-attention shape of `attention` (32 queries, d = 64, heads % 4 == 0,
    64 < Sk <= 257) runs the tcgen05 kernel; False = the generic mma.sync kernel (cross-check)."""
    check(_lib.load().tgb_attention_set_tc(1 if on else 0))


def vit_attention_set_variant(variant: int) -> None:
    """Test knob: 0 = ping-pong kernel (default), 1 = one-shot kernel."""
    check(_lib.load().tgb_vit_attention_set_variant(int(variant)))


def vit_attention(qkv: torch.Tensor, frames: int, scale: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """tcgen05 self-attention for the EVA-ViT-g shape: qkv bf16 [frames*257, 4224] -> ctx bf16 [frames*257, 1408]."""
    _need(qkv, torch.bfloat16, "qkv", contiguous=False)
    if qkv.dim() != 2 or qkv.stride(1) != 1 or qkv.shape[0] != frames * 257 or qkv.shape[1] < 4224:
        raise _lib.TgbError("vit_attention: qkv must be [frames*257, >=4224] bf16")
    if out is None:
        out = torch.empty((frames * 257, 1408), dtype=torch.bfloat16, device=qkv.device)
    global _PROFILE
    if _PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(_lib.load().tgb_vit_attention_bf16(_p(qkv), qkv.stride(0), _p(out), out.stride(0), frames, float(scale), _stream()))
    if _PROFILE is not None:
        e1.record()
        _PROFILE.append(("attention", (frames, 16, 257, 257, 88), e0, e1))
    return out


# ------------------------------------------------------------------------------------------------ N4
def llm_inputs(visual: torch.Tensor, table: torch.Tensor, ids: torch.Tensor, qmask: torch.Tensor):
    """visual f32 [B,Nv,H], table f32 [V,H], ids/qmask i64 [B,Tq] -> (inputs_embeds f32 [B,Nv+Tq,H], attention_mask i64)."""
    v = _need(visual, torch.float32, "visual")
    tb = _need(table, torch.float32, "table")
    i = _need(ids, torch.int64, "ids")
    qm = _need(qmask, torch.int64, "qmask")
    if v.dim() != 3 or tb.dim() != 2 or tb.shape[1] != v.shape[2] or i.shape != qm.shape or i.dim() != 2 or i.shape[0] != v.shape[0]:
        raise _lib.TgbError("visual [B,Nv,H], table [V,H], ids/qmask [B,Tq] expected")
    b, nv, h = v.shape
    tq = i.shape[1]
    emb = torch.empty((b, nv + tq, h), dtype=torch.float32, device=v.device)
    mask = torch.empty((b, nv + tq), dtype=torch.int64, device=v.device)
    check(_lib.load().tgb_llm_inputs_f32(_p(v), _p(tb), _p(i), _p(qm), tb.shape[0], b, nv, tq, h, _p(emb), _p(mask), _stream()))
    return emb, mask


def replace_ids_(ids: torch.Tensor, src: int, dst: int) -> torch.Tensor:
    i = _need(ids, torch.int64, "ids")
    check(_lib.load().tgb_replace_ids(_p(i), i.numel(), int(src), int(dst), _stream()))
    return ids


# ------------------------------------------------------------------------------------------------ N1 (RAFT)
def _ld(t: torch.Tensor) -> int:
    """Row pitch (elements) of a 2-D view whose last dimension is contiguous."""
    if t.dim() != 2 or t.stride(1) != 1:
        raise _lib.TgbError("expected a 2-D tensor with a contiguous last dimension")
    return t.stride(0)


def im2col_nhwc(x: torch.Tensor, n: int, h: int, w: int, c: int, kh: int, kw: int, stride=(1, 1), padding=(0, 0),
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x bf16 pixel rows [n*h*w, >=c] -> bf16 [n*ho*wo, roundup8(kh*kw*c)]."""
    _need(x, torch.bfloat16, "x", contiguous=False)
    ho = (h + 2 * padding[0] - kh) // stride[0] + 1
    wo = (w + 2 * padding[1] - kw) // stride[1] + 1
    kpad = (kh * kw * c + 7) // 8 * 8
    if out is None:
        out = torch.empty((n * ho * wo, kpad), dtype=torch.bfloat16, device=x.device)
    check(_lib.load().tgb_im2col_nhwc_bf16(_p(x), _ld(x), n, h, w, c, kh, kw, stride[0], stride[1], padding[0], padding[1],
                                           _p(out), _ld(out), _stream()))
    return out


def instance_norm_nhwc(x: torch.Tensor, n: int, hw: int, relu: bool, want_f32: bool, want_bf16: bool, eps: float = 1e-5):
    x = _need(x, torch.float32, "x")
    c = x.shape[-1]
    of = torch.empty_like(x) if want_f32 else None
    ob = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device) if want_bf16 else None
    ws = torch.empty((n * 17 * c * 2,), dtype=torch.float32, device=x.device)
    check(_lib.load().tgb_instance_norm_nhwc(_p(x), n, hw, c, float(eps), int(relu), _p(ws), _p(of), _p(ob), _stream()))
    return of, ob


ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH = range(4)


def act(x: torch.Tensor, mode: int, y: Optional[torch.Tensor] = None, scale: float = 1.0, out_f32: Optional[torch.Tensor] = None,
        out_bf16: Optional[torch.Tensor] = None, want_f32: bool = False, want_bf16: bool = False):
    """act(scale*x (+ y)) on 2-D fp32 views (row pitches allowed) -> (f32 or None, bf16 or None)."""
    _need(x, torch.float32, "x", contiguous=False)
    m, n = x.shape
    if out_f32 is None and want_f32:
        out_f32 = torch.empty((m, n), dtype=torch.float32, device=x.device)
    if out_bf16 is None and want_bf16:
        out_bf16 = torch.empty((m, n), dtype=torch.bfloat16, device=x.device)
    check(_lib.load().tgb_act_f32(_p(x), _ld(x), _p(y), _ld(y) if y is not None else 0, m, n, int(mode), float(scale),
                                  _p(out_f32), _ld(out_f32) if out_f32 is not None else 0,
                                  _p(out_bf16), _ld(out_bf16) if out_bf16 is not None else 0, _stream()))
    return out_f32, out_bf16


def gru_rh(zr: torch.Tensor, h: torch.Tensor, rh_out: torch.Tensor) -> None:
    _need(zr, torch.float32, "zr"); _need(h, torch.float32, "h"); _need(rh_out, torch.bfloat16, "rh", contiguous=False)
    check(_lib.load().tgb_gru_rh(_p(zr), _p(h), h.shape[0], h.shape[1], _p(rh_out), _ld(rh_out), _stream()))


def gru_update(zr: torch.Tensor, q: torch.Tensor, h: torch.Tensor, h_bf16_out: torch.Tensor) -> None:
    _need(zr, torch.float32, "zr"); _need(q, torch.float32, "q"); _need(h, torch.float32, "h")
    _need(h_bf16_out, torch.bfloat16, "h_bf16", contiguous=False)
    check(_lib.load().tgb_gru_update(_p(zr), _p(q), _p(h), h.shape[0], h.shape[1], _p(h_bf16_out), _ld(h_bf16_out), _stream()))


def corr_pool(x: torch.Tensor) -> torch.Tensor:
    x = _need(x, torch.float32, "corr")
    r, h, w = x.shape
    out = torch.empty((r, h // 2, w // 2), dtype=torch.float32, device=x.device)
    check(_lib.load().tgb_corr_pool(_p(x), _p(out), r, h, w, _stream()))
    return out


def corr_lookup(pyr, flow: torch.Tensor, n: int, h: int, w: int, out: torch.Tensor) -> torch.Tensor:
    for t in pyr:
        _need(t, torch.float32, "pyramid level")
    _need(flow, torch.float32, "flow")
    _need(out, torch.bfloat16, "out", contiguous=False)
    check(_lib.load().tgb_corr_lookup(_p(pyr[0]), _p(pyr[1]), _p(pyr[2]), _p(pyr[3]), _p(flow), n, h, w, _p(out), _ld(out), _stream()))
    return out


def raft_upsample(mask: torch.Tensor, flow: torch.Tensor, n: int, h: int, w: int) -> torch.Tensor:
    _need(mask, torch.float32, "mask", contiguous=False)
    _need(flow, torch.float32, "flow")
    out = torch.empty((n, 2, 8 * h, 8 * w), dtype=torch.float32, device=mask.device)
    check(_lib.load().tgb_raft_upsample(_p(mask), _ld(mask), _p(flow), n, h, w, _p(out), _stream()))
    return out


def raft_prep_image(img: torch.Tensor, ldo: int = 8) -> torch.Tensor:
    img = _need(img, torch.float32, "image")
    n, c, h, w = img.shape
    out = torch.zeros((n * h * w, ldo), dtype=torch.bfloat16, device=img.device)
    check(_lib.load().tgb_raft_prep_image(_p(img), n, c, h * w, _p(out), ldo, _stream()))
    return out
