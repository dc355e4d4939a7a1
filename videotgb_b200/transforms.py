"""Frame normalisation behind the reference's config keys.

The reference composes ResizeVideo -> ToUint8 -> ToTHWC -> ToTensorVideo -> NormalizeVideo on the CPU dataloader
(src/data/videoinstruct_datamodule.py:125-132, src/gadgets/transforms.py).  `VideoNormalizer` runs the
whole chain on the GPU (decoded u8 THWC at native resolution -> fp32 TCHW at target_size, bit-exact); `processor_name`, `nframe` and
`target_size` keep their reference meaning (configs/data/videoinstruct.yaml:5-9).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import ops


@dataclass
class DataConfig:
    processor_name: str = "Salesforce/blip2-flan-t5-xl"
    sampler_processor_name: str = "bert-base-uncased"
    nframe: int = 4
    target_size: int = 224


class VideoNormalizer:
    mean = ops.CLIP_MEAN
    std = ops.CLIP_STD

    def __init__(self, config: DataConfig = DataConfig()):
        self.config = config

    def resize(self, frames_u8_thwc: torch.Tensor) -> torch.Tensor:
        """ResizeVideo -> ToUint8 -> ToTHWC: decoded u8 [T,H0,W0,3] -> u8 [T,S,S,3], S = target_size."""
        s = self.config.target_size
        if frames_u8_thwc.shape[1] == s and frames_u8_thwc.shape[2] == s:
            return frames_u8_thwc                        # the interpolation is the identity at equal sizes
        return ops.resize_bilinear_u8(frames_u8_thwc.contiguous(), s)

    def __call__(self, frames_u8_thwc: torch.Tensor) -> torch.Tensor:
        """Decoded u8 [T,H0,W0,3] on the GPU -> fp32 [T,3,S,S]: the whole reference transform chain."""
        return ops.normalize_frames_u8(self.resize(frames_u8_thwc), self.mean, self.std)
