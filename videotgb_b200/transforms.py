"""Frame normalisation behind the reference's config keys.

The reference composes ResizeVideo -> ToUint8 -> ToTHWC -> ToTensorVideo -> NormalizeVideo on the CPU dataloader
(src/data/videoinstruct_datamodule.py:125-132, src/gadgets/transforms.py).  `VideoNormalizer` runs the
ToTensorVideo+NormalizeVideo part (u8 THWC -> fp32 TCHW, bit-exact) on the GPU; `processor_name`, `nframe` and
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

    def __call__(self, frames_u8_thwc: torch.Tensor) -> torch.Tensor:
        """u8 [T,H,W,3] on the GPU -> fp32 [T,3,H,W]."""
        t, h, w, _ = frames_u8_thwc.shape
        if h != self.config.target_size or w != self.config.target_size:
            raise ValueError(f"frames must already be {self.config.target_size}x{self.config.target_size}; the bilinear "
                             "ResizeVideo stage is a 'next' row (SURVEY.md section 8f N3)")
        return ops.normalize_frames_u8(frames_u8_thwc, self.mean, self.std)
