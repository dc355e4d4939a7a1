"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference video transform.

Follows ``src/gadgets/transforms.py`` ``ToUint8`` :213-221, ``ToTHWC`` :224-239, ``ToTensorVideo``
:167-186, ``NormalizeVideo`` :142-164 and ``src/gadgets/functional_video.py`` ``to_tensor`` :76-90,
``normalize`` :93-109, composed as in ``src/data/videoinstruct_datamodule.py:125-132`` followed by the
caller's ``permute(1,0,2,3)``.  The bilinear ``ResizeVideo`` (:33-60) is the identity for the synthetic
224x224 inputs and is restated with ``torch.nn.functional.interpolate`` for other sizes.

Exact op order (each op rounds to fp32):  x = float32(u8) / 255.0 ; x = x - mean_c ; x = x / std_c
"""
import numpy as np

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def normalize_u8_thwc(frames_u8: np.ndarray, mean=CLIP_MEAN, std=CLIP_STD) -> np.ndarray:
    """u8 [T,H,W,3] -> fp32 [T,3,H,W]."""
    assert frames_u8.dtype == np.uint8 and frames_u8.shape[-1] == 3
    x = frames_u8.astype(np.float32) / np.float32(255.0)
    x = x - np.asarray(mean, dtype=np.float32)
    x = x / np.asarray(std, dtype=np.float32)
    return np.ascontiguousarray(x.transpose(0, 3, 1, 2))


def resize_to_uint8(clip_cthw, size: int):
    """float [3,T,H0,W0] -> u8 [T,size,size,3]: bilinear(align_corners=False) then truncation to uint8."""
    import torch
    import torch.nn.functional as F
    c = torch.as_tensor(clip_cthw, dtype=torch.float32)
    r = F.interpolate(c, size=(size, size), mode="bilinear", align_corners=False)
    return r.to(torch.uint8).permute(1, 2, 3, 0).contiguous().numpy()


# --------------------------------------------------------------------------------------------
# ResizeVideo -> ToUint8 -> ToTHWC restated in numpy
#   src/gadgets/functional_video.py:33-40 (F.interpolate(clip, size, mode="bilinear", align_corners=False), no antialias)
#   src/gadgets/transforms.py:213-220 (ToUint8: truncation), :223-238 (ToTHWC)
# F.interpolate's CPU arithmetic is not part of the reference repo (it is ATen's upsample_bilinear2d, torch 2.11 as
# installed in this image); the exact fp32 operation order below -- source index through ONE fused multiply-add,
# horizontal pass first, each pass evaluated as fma(x0, w0, round(x1*w1)) -- was determined against F.interpolate
# bit for bit (tests/test_oracle_golden.py pins it on the reference-generated fixture and on live F.interpolate).
# --------------------------------------------------------------------------------------------
def _fma32(a, b, c):
    """fp32 fused multiply-add of float32 arrays: the product of two float32 values is exact in float64 and the one
    float64 addition is then rounded to float32 (double rounding cannot occur for 24-bit x 24-bit + 24-bit here except
    in astronomically rare ties; the fixture comparison would reveal it)."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def bilinear_tables(in_size: int, out_size: int):
    """ATen `compute_source_index_and_lambda` (align_corners=False): -> (idx0, idx1, lambda0, lambda1)."""
    f32 = np.float32
    i = np.arange(out_size, dtype=np.float32)
    if in_size == out_size:
        idx = np.arange(out_size, dtype=np.int64)
        return idx, idx, np.ones(out_size, f32), np.zeros(out_size, f32)
    scale = f32(in_size) / f32(out_size)
    real = _fma32(np.full(out_size, scale, f32), i + f32(0.5), np.full(out_size, -0.5, f32))
    real = np.maximum(real, f32(0))
    idx0 = np.minimum(real.astype(np.int64), in_size - 1)
    lam1 = np.clip(real - idx0.astype(np.float32), f32(0), f32(1)).astype(np.float32)
    idx1 = idx0 + (idx0 < in_size - 1)
    return idx0, idx1, (f32(1) - lam1).astype(np.float32), lam1


def resize_bilinear_u8(frames_thwc_u8: np.ndarray, size: int) -> np.ndarray:
    """u8 [T,H0,W0,3] (decoded RGB, what `.float()` of the reference pipeline holds) -> u8 [T,size,size,3]."""
    assert frames_thwc_u8.dtype == np.uint8 and frames_thwc_u8.shape[-1] == 3
    t, h, w, _ = frames_thwc_u8.shape
    h0, h1, a0, a1 = bilinear_tables(h, size)
    w0, w1, b0, b1 = bilinear_tables(w, size)
    s = frames_thwc_u8.astype(np.float32)
    b0, b1 = b0[None, None, :, None], b1[None, None, :, None]
    a0, a1 = a0[None, :, None, None], a1[None, :, None, None]

    def row(hh):                                                 # horizontal pass on source rows hh -> [T,size,size,3]
        r = s[:, hh]
        return _fma32(r[:, :, w0], b0, (r[:, :, w1] * b1).astype(np.float32))
    out = _fma32(row(h0), a0, (row(h1) * a1).astype(np.float32))
    return out.astype(np.uint8)                                  # values lie in [0,255]: truncation like .to(torch.uint8)
