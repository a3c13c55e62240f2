"""CPU oracle for the paste-back step that follows the generator  --  TEST INFRASTRUCTURE ONLY.

Restates, in numpy integer / float32 arithmetic, what the reference computes with OpenCV:

    prepare_paste_back(mask_crop, crop_M_c2o, dsize, if_float=True)   reference src/utils/crop.py:515-521
    paste_back(img_crop, M_c2o, img_ori, mask_ori)                    reference src/utils/crop.py:523-529
    _transform_img -> cv2.warpAffine(img, M[:2, :], dsize, flags=cv2.INTER_LINEAR)   reference src/utils/crop.py:49-63

The arithmetic lives in a third-party dependency of the reference (opencv-python, un-pinned in requirements.txt; 4.13.0 in
this image).  Its published algorithm (modules/imgproc/src/imgwarp.cpp: warpAffine + remapBilinear), restated here:

  * the 2x3 matrix is inverted in double precision (warpAffine without WARP_INVERSE_MAP);
  * source coordinates are 10-bit fixed point: X = (round((M01*y + M02) * 1024) + 16 + round(M00 * x * 1024)) >> 5, same for Y
    (round = round-half-to-even of a double, AB_BITS = 10, INTER_BITS = 5, round_delta = 16);  sx = X >> 5, fx = X & 31;
  * uint8 images: weights 32*(32-fy)*(32-fx) ... (exact, they sum to 2^15: the table fix-up never triggers for bilinear),
    value = (sum w*S + 2^14) >> 15;  float32 images: float weights (1-b)*(1-a) ... with a = fx/32, value =
    ((S00*w0 + S01*w1) + S10*w2) + S11*w3 in float32;  taps outside the source read the border value 0 (BORDER_CONSTANT);
  * paste_back: clip(mask*result + (1-mask)*img_ori, 0, 255).astype(uint8) in float32 (numpy promotion), truncating.

PARITY PINNING: tests/test_pasteback.py checks every function bit for bit against cv2 itself (present in this image) and,
when /root/reference is present, against the reference's own prepare_paste_back / paste_back; golden vectors made by those
reference functions are committed under tests/golden/pasteback_*.npz (tests/golden/make_golden.py pasteback).
"""
from __future__ import annotations

import numpy as np


def invert_affine(M) -> np.ndarray:
    """cv::warpAffine's inversion of the 2x3 matrix (imgwarp.cpp), double precision, same operation order."""
    M = np.asarray(M, dtype=np.float64)[:2, :3].copy()
    D = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = M[1, 1] * D, M[0, 0] * D
    iM = np.zeros((2, 3), dtype=np.float64)
    iM[0, 0] = A11
    iM[0, 1] = M[0, 1] * (-D)
    iM[1, 0] = M[1, 0] * (-D)
    iM[1, 1] = A22
    iM[0, 2] = -iM[0, 0] * M[0, 2] - iM[0, 1] * M[1, 2]
    iM[1, 2] = -iM[1, 0] * M[0, 2] - iM[1, 1] * M[1, 2]
    return iM


def source_coords(iM: np.ndarray, W: int, H: int):
    """Integer source position (sx, sy) and 5-bit fractions (fx, fy) of every destination pixel, [H, W] int64."""
    xs = np.arange(W, dtype=np.float64)
    ys = np.arange(H, dtype=np.float64)
    adelta = np.rint(iM[0, 0] * xs * 1024.0).astype(np.int64)
    bdelta = np.rint(iM[1, 0] * xs * 1024.0).astype(np.int64)
    X0 = np.rint((iM[0, 1] * ys + iM[0, 2]) * 1024.0).astype(np.int64) + 16
    Y0 = np.rint((iM[1, 1] * ys + iM[1, 2]) * 1024.0).astype(np.int64) + 16
    X = (X0[:, None] + adelta[None, :]) >> 5
    Y = (Y0[:, None] + bdelta[None, :]) >> 5
    return X >> 5, Y >> 5, X & 31, Y & 31


def _taps(img: np.ndarray, sx, sy):
    """The 2x2 neighbourhood of every destination pixel with BORDER_CONSTANT 0; img [h, w, c]."""
    h, w = img.shape[:2]
    out = []
    for dy, dx in ((0, 0), (0, 1), (1, 0), (1, 1)):
        yy, xx = sy + dy, sx + dx
        ok = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
        v = img[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)]
        out.append(np.where(ok[..., None], v, img.dtype.type(0)))
    return out


def warp_affine_u8(img: np.ndarray, M, dsize) -> np.ndarray:
    """cv2.warpAffine(img uint8 [h,w,c], M, (W, H), flags=INTER_LINEAR)."""
    W, H = int(dsize[0]), int(dsize[1])
    sx, sy, fx, fy = source_coords(invert_affine(M), W, H)
    s = [t.astype(np.int64) for t in _taps(img, sx, sy)]
    w = [32 * (32 - fy) * (32 - fx), 32 * (32 - fy) * fx, 32 * fy * (32 - fx), 32 * fy * fx]
    acc = sum(wk[..., None] * sk for wk, sk in zip(w, s))
    return np.clip((acc + (1 << 14)) >> 15, 0, 255).astype(np.uint8)


def warp_affine_f32(img: np.ndarray, M, dsize) -> np.ndarray:
    """cv2.warpAffine(img float32 [h,w,c], M, (W, H), flags=INTER_LINEAR)."""
    W, H = int(dsize[0]), int(dsize[1])
    sx, sy, fx, fy = source_coords(invert_affine(M), W, H)
    one = np.float32(1)
    a = fx.astype(np.float32) * np.float32(1.0 / 32.0)
    b = fy.astype(np.float32) * np.float32(1.0 / 32.0)
    w = [(one - b) * (one - a), (one - b) * a, b * (one - a), b * a]
    s = _taps(img.astype(np.float32, copy=False), sx, sy)
    r = s[0] * w[0][..., None]
    r = r + s[1] * w[1][..., None]
    r = r + s[2] * w[2][..., None]
    r = r + s[3] * w[3][..., None]
    return r.astype(np.float32)


def prepare_paste_back(mask_crop: np.ndarray, crop_M_c2o, dsize, if_float: bool = False) -> np.ndarray:
    """reference src/utils/crop.py:515-521."""
    if mask_crop.dtype == np.uint8:
        mask_ori = warp_affine_u8(mask_crop, crop_M_c2o, dsize)
    else:
        mask_ori = warp_affine_f32(mask_crop, crop_M_c2o, dsize)
    if if_float is False:
        mask_ori = mask_ori.astype(np.float32) / 255.
    return mask_ori


def paste_back(img_crop: np.ndarray, M_c2o, img_ori: np.ndarray, mask_ori: np.ndarray) -> np.ndarray:
    """reference src/utils/crop.py:523-529."""
    dsize = (img_ori.shape[1], img_ori.shape[0])
    result = warp_affine_u8(img_crop, M_c2o, dsize)
    return np.clip(mask_ori * result + (1 - mask_ori) * img_ori, 0, 255).astype(np.uint8)


def paste_back_frame(img_crop: np.ndarray, mask_crop: np.ndarray, M_c2o, img_ori: np.ndarray) -> np.ndarray:
    """The two calls of the per-frame loop, reference src/can_swap_pipeline_e2e.py:277-282: mask_crop [h,w] or [h,w,3] float32."""
    if mask_crop.ndim == 2:
        mask_crop = np.stack([mask_crop] * 3, axis=-1)
    mask_ori = prepare_paste_back(mask_crop, M_c2o, dsize=(img_ori.shape[1], img_ori.shape[0]), if_float=True)
    return paste_back(img_crop, M_c2o, img_ori, mask_ori)


def soft_erosion(x, kernel_size=21, threshold=0.9, iterations=3):
    """SoftErosion (reference src/utils/crop.py:21-47) in plain torch on the CPU: x [B,1,H,W] -> (soft mask, x >= threshold).
    Float work: compared with a tolerance, away from the threshold (tests/test_pasteback.py)."""
    import torch
    import torch.nn.functional as F
    r = kernel_size // 2
    yi, xi = torch.meshgrid(torch.arange(0., kernel_size), torch.arange(0., kernel_size), indexing="ij")
    dist = torch.sqrt((xi - r) ** 2 + (yi - r) ** 2)
    kernel = dist.max() - dist
    kernel /= kernel.sum()
    weight = kernel.view(1, 1, *kernel.shape)
    x = x.float().clone()
    for _ in range(iterations - 1):
        x = torch.min(x, F.conv2d(x, weight=weight, groups=x.shape[1], padding=r))
    x = F.conv2d(x, weight=weight, groups=x.shape[1], padding=r)
    pre = x.clone()
    mask = x >= threshold
    x[mask] = 1.0
    x[~mask] /= x[~mask].max()
    return x, mask, pre
