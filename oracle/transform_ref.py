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
