"""Host-side mirror of the paste-back step that follows the generator.

    SoftErosion                      reference src/utils/crop.py:21-47 (same constructor, same `weight` buffer, forward -> (x, mask))
    prepare_paste_back / paste_back  reference src/utils/crop.py:515-529, fused: `paste_back_frames`

The reference moves the mask to the host and runs two full-frame `cv2.warpAffine` calls plus a float blend on the CPU per
frame (src/can_swap_pipeline_e2e.py:274-283); here the tensors stay on the device and one kernel does the work, bit-exact
with OpenCV's arithmetic (oracle/pasteback_oracle.py).  No arithmetic lives in this file.
"""
from __future__ import annotations

import torch
from torch import nn

from .engine import Engine


class SoftErosion(nn.Module):
    def __init__(self, kernel_size=15, threshold=0.6, iterations=1, engine: Engine = None):
        super().__init__()
        r = kernel_size // 2
        self.padding = r
        self.iterations = iterations
        self.threshold = threshold
        # the weight buffer exactly as the reference builds it (crop.py:29-35)
        y_indices, x_indices = torch.meshgrid(torch.arange(0., kernel_size), torch.arange(0., kernel_size), indexing="ij")
        dist = torch.sqrt((x_indices - r) ** 2 + (y_indices - r) ** 2)
        kernel = dist.max() - dist
        kernel /= kernel.sum()
        self.register_buffer("weight", kernel.view(1, 1, *kernel.shape))
        self._engine = engine

    def bind(self, engine: Engine):
        self._engine = engine
        return self

    def forward(self, x: torch.Tensor):
        """x [B,1,H,W] (any float dtype) on the engine's device -> (soft mask [B,1,H,W] f32, x >= threshold)."""
        if self._engine is None:
            raise RuntimeError("SoftErosion: bind(engine) first (there is no CPU path)")
        B, C, H, W = x.shape
        if C != 1:
            raise ValueError("SoftErosion: single-channel masks (the pipeline calls it with [1,1,H,W])")
        w = self.weight.to(x.device)
        out, hard = self._engine.soft_erosion(x.float().reshape(B, H, W), w[0, 0], self.threshold, self.iterations)
        return out.reshape(B, 1, H, W), hard.reshape(B, 1, H, W)


def paste_back_frames(engine: Engine, img_crop: torch.Tensor, mask_crop: torch.Tensor, M_c2o, img_ori: torch.Tensor, out=None):
    """mask_ori = prepare_paste_back(stack3(mask_crop), M_c2o, dsize, if_float=True); paste_back(img_crop, M_c2o, img_ori, mask_ori)
    for a batch of frames, on the device (reference src/can_swap_pipeline_e2e.py:277-282)."""
    return engine.paste_back(img_crop, mask_crop, M_c2o, img_ori, out=out)
