"""Engine: one `cs_ctx` of libcanonswap_b200.so bound to a CUDA device.

Thin and explicit: torch is used only for device memory (`tensor.data_ptr()`), the current CUDA
stream and lifetime; every computation is a C-ABI call into the hand-written sm_100a kernels.
No call here ever falls back to PyTorch ops.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Mapping, Optional

import torch

from . import _lib, spec


class CanonSwapError(RuntimeError):
    """A negative cs_status from the C ABI (message = cs_last_error)."""


def _flat_table(weights: Mapping[str, Mapping[str, torch.Tensor]]):
    """`combined_weights.pth`-layout dict -> (cs_tensor_desc array, keep-alive list)."""
    names, tensors = [], []
    for net in spec.NETS + (spec.MOTION_NET,):
        if net not in weights:
            if net == spec.MOTION_NET:        # optional: the path takes the keypoints as inputs without it
                continue
            raise KeyError(f"weights dict lacks the '{net}' state_dict (reference can_swap_e2e.py:93-98)")
        for key, t in weights[net].items():
            t = t.detach()
            if t.dtype == torch.float32:
                dt = _lib.CS_F32
            elif t.dtype == torch.int64:
                dt = _lib.CS_I64
            else:
                raise TypeError(f"{net}.{key}: unsupported dtype {t.dtype} (the reference checkpoint is fp32)")
            names.append((f"{net}.{key}".encode(), dt))
            tensors.append(t.to("cpu").contiguous())
    arr = (_lib.TensorDesc * len(names))()
    for i, ((name, dt), t) in enumerate(zip(names, tensors)):
        arr[i].name = name
        arr[i].data = t.data_ptr()
        arr[i].dtype = dt
        arr[i].ndim = t.dim()
        for j, s in enumerate(t.shape):
            arr[i].shape[j] = s
    return arr, (names, tensors)


class Engine:
    """Owns a cs_ctx: packed weights + workspace for frames of `net_hw` up to `max_batch` per call."""

    def __init__(self, weights: Mapping[str, Mapping[str, torch.Tensor]], net_hw=(256, 256), max_batch: int = 8,
                 device: int = 0, conv_impl: int = 0, options: Mapping[int, int] | None = None):
        self._lib = _lib.load()
        self._ctx = C.c_void_p()
        self.device = torch.device("cuda", device)
        self.net_h, self.net_w = int(net_hw[0]), int(net_hw[1])
        self.h, self.w = self.net_h // 4, self.net_w // 4
        self.max_batch = int(max_batch)
        rc = self._lib.cs_create(C.byref(self._ctx), device, self.max_batch, self.net_h, self.net_w)
        if rc != 0:
            msg = self._lib.cs_last_error(None).decode()
            self._ctx = C.c_void_p()
            raise CanonSwapError(f"cs_create failed ({rc}): {msg}")
        if conv_impl:
            self.set_option(_lib.CS_OPT_CONV_IMPL, conv_impl)
        for k, v in (options or {}).items():      # options that shape the weight packing must precede cs_load_weights
            self.set_option(int(k), int(v))
        if weights is not None:          # None: kernel-level test entry points only (no networks)
            table, keep = _flat_table(weights)
            self._check(self._lib.cs_load_weights(self._ctx, table, len(table)))
            del keep
        self._identity = None

    # ------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self._lib.cs_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _live(self):
        if not self._ctx.value:
            raise CanonSwapError("this Engine was closed (its weights were reloaded or a larger batch was requested): "
                                 "fetch the current one from can_swapper.engine(...)")
        return self._ctx

    def _check(self, rc: int):
        if rc != 0 and not self._ctx.value:
            self._live()
        if rc != 0:
            raise CanonSwapError(f"canonswap_b200 error {rc}: {self._lib.cs_last_error(self._ctx).decode()}")

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _in(self, t: torch.Tensor, shape, dtype=torch.float32, name="tensor", output: bool = False) -> torch.Tensor:
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"{name}: expected a torch.Tensor, got {type(t)}")
        if t.device != self.device:
            raise ValueError(f"{name}: tensor is on {t.device}, engine is on {self.device} (no implicit copies)")
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"{name}: shape {tuple(t.shape)} != expected {tuple(shape)}")
        if t.dtype != dtype:
            raise TypeError(f"{name}: dtype {t.dtype} != expected {dtype}")
        if output and not t.is_contiguous():
            raise ValueError(f"{name}: output buffers must be contiguous (a temporary copy would never reach the caller)")
        return t.detach().contiguous()

    def _batch(self, t: torch.Tensor) -> int:
        B = int(t.shape[0])
        if not 1 <= B <= self.max_batch:
            raise ValueError(f"batch {B} outside [1, max_batch={self.max_batch}]")
        return B

    def _new(self, *shape, dtype=torch.float32):
        return torch.empty(shape, dtype=dtype, device=self.device)

    # ------------------------------------------------------------------------------------------
    def set_option(self, option: int, value: int):
        self._check(self._lib.cs_set_option(self._ctx, option, value))

    @property
    def launch_count(self) -> int:
        return int(self._lib.cs_launch_count(self._ctx))

    @property
    def workspace_bytes(self) -> int:
        return int(self._lib.cs_workspace_bytes(self._ctx))

    def set_identity(self, source_id: torch.Tensor):
        """source_id [1,512] (or [512]) fp32 on the engine device: ArcFace identity, L2-normalised
        (reference can_swap_e2e.py:102-107)."""
        sid = source_id.reshape(-1)
        sid = self._in(sid, (spec.LATENT,), name="source_id")
        self._check(self._lib.cs_set_identity(self._ctx, sid.data_ptr(), self._stream()))
        self._identity = sid

    # ---- per-stage calls (reference module forwards) ------------------------------------------
    def appearance(self, x: torch.Tensor) -> torch.Tensor:
        B = self._batch(x)
        x = self._in(x, (B, 3, self.net_h, self.net_w), name="x")
        out = self._new(B, 32, 16, self.h, self.w)
        self._check(self._lib.cs_appearance(self._ctx, x.data_ptr(), out.data_ptr(), B, self._stream()))
        return out

    def _kp(self, kp, B, name):
        return self._in(kp, (B, spec.NUM_KP, 3), name=name)

    def warp(self, feature_3d, kp_source, kp_driving, want_deformation=False):
        B = self._batch(feature_3d)
        f = self._in(feature_3d, (B, 32, 16, self.h, self.w), name="feature_3d")
        ks, kd = self._kp(kp_source, B, "kp_source"), self._kp(kp_driving, B, "kp_driving")
        out = self._new(B, 32, 16, self.h, self.w)
        occ = self._new(B, 1, self.h, self.w)
        deform = self._new(B, 16, self.h, self.w, 3) if want_deformation else None
        self._check(self._lib.cs_warp(self._ctx, f.data_ptr(), ks.data_ptr(), kd.data_ptr(), out.data_ptr(),
                                      occ.data_ptr(), deform.data_ptr() if deform is not None else None, B,
                                      self._stream()))
        return (out, occ, deform) if want_deformation else (out, occ)

    def warp_out(self, out3d, occlusion_map=None):
        B = self._batch(out3d)
        f = self._in(out3d, (B, 32, 16, self.h, self.w), name="out")
        occ = self._in(occlusion_map, (B, 1, self.h, self.w), name="occlusion_map") if occlusion_map is not None else None
        out = self._new(B, 256, self.h, self.w)
        self._check(self._lib.cs_warp_out(self._ctx, f.data_ptr(), occ.data_ptr() if occ is not None else None,
                                          out.data_ptr(), B, self._stream()))
        return out

    def warp_forward(self, feature_3d, kp_driving, kp_source):
        B = self._batch(feature_3d)
        f = self._in(feature_3d, (B, 32, 16, self.h, self.w), name="feature_3d")
        ks, kd = self._kp(kp_source, B, "kp_source"), self._kp(kp_driving, B, "kp_driving")
        out = self._new(B, 256, self.h, self.w)
        occ = self._new(B, 1, self.h, self.w)
        deform = self._new(B, 16, self.h, self.w, 3)
        self._check(self._lib.cs_warp_forward(self._ctx, f.data_ptr(), kd.data_ptr(), ks.data_ptr(), out.data_ptr(),
                                              occ.data_ptr(), deform.data_ptr(), B, self._stream()))
        return {"occlusion_map": occ, "deformation": deform, "out": out}

    def swap(self, feature_3d, return_mask=False):
        if self._identity is None:
            raise CanonSwapError("swap: no identity set (call set_identity first)")
        B = self._batch(feature_3d)
        f = self._in(feature_3d, (B, 32, 16, self.h, self.w), name="x")
        out = self._new(B, 32, 16, self.h, self.w)
        masks = self._new(7, B, 1, self.h, self.w) if return_mask else None
        self._check(self._lib.cs_swap(self._ctx, f.data_ptr(), out.data_ptr(),
                                      masks.data_ptr() if masks is not None else None, B, self._stream()))
        return (out, list(masks.unbind(0))) if return_mask else out

    def refine(self, feature_3d):
        B = self._batch(feature_3d)
        f = self._in(feature_3d, (B, 32, 16, self.h, self.w), name="x")
        out = self._new(B, 32, 16, self.h, self.w)
        self._check(self._lib.cs_refine(self._ctx, f.data_ptr(), out.data_ptr(), B, self._stream()))
        return out

    def spade(self, feature, want_u8=False):
        B = self._batch(feature)
        f = self._in(feature, (B, 256, self.h, self.w), name="feature")
        img = self._new(B, 3, 2 * self.net_h, 2 * self.net_w)
        u8 = self._new(B, 2 * self.net_h, 2 * self.net_w, 3, dtype=torch.uint8) if want_u8 else None
        self._check(self._lib.cs_spade(self._ctx, f.data_ptr(), img.data_ptr(),
                                       u8.data_ptr() if u8 is not None else None, B, self._stream()))
        return (img, u8) if want_u8 else img

    # ---- the fused loop body --------------------------------------------------------------------
    def frame(self, frames: torch.Tensor, kp_t: Optional[torch.Tensor] = None, kp_can: Optional[torch.Tensor] = None,
              out_u8: Optional[torch.Tensor] = None, out_f32: Optional[torch.Tensor] = None, debug_decodes: bool = False,
              v2i: bool = False, motion: bool = False, v2i_feature: bool = False, batch: Optional[int] = None):
        """One batch of the per-frame loop body (reference can_swap_pipeline_e2e.py:242-267).

        frames: [B,net_h,net_w,3] uint8 (HWC, as cropped) or [B,3,net_h,net_w] fp32 in [0,1]
        (v2i_feature=True: ONE appearance volume [1,32,16,h,w] for the whole batch, see CS_FRAME_V2I_FEATURE);
        kp_t = x_t_info['x_s'], kp_can = scale * kp  (both [B,21,3]); with motion=True they are derived on the device
        from the frames by the motion extractor (reference can_swap_pipeline_e2e.py:112-125,231-243) and may be None.
        Returns (out_u8 [B,2H,2W,3] uint8, out_f32 [B,3,2H,2W] or None).
        """
        v2i = v2i or v2i_feature
        if self._identity is None and not v2i:
            raise CanonSwapError("frame: no identity set (call set_identity first)")
        B = self._batch(kp_t) if v2i_feature else self._batch(frames)
        flags = _lib.CS_FRAME_DEBUG_DECODES if debug_decodes else 0
        if v2i:      # reference can_swap_pipeline_v2i.py:308-309: kp_t = kp_source, kp_can = kp_driving, no swap / refine
            flags |= _lib.CS_FRAME_V2I
        if v2i_feature:   # `frames` = ONE appearance volume [1,32,16,h,w] shared by the B keypoint sets (CS_FRAME_V2I_FEATURE)
            flags |= _lib.CS_FRAME_V2I_FEATURE
            fr = self._in(frames, (1, 32, 16, self.h, self.w), name="feature")
        elif frames.dtype == torch.uint8:
            fr = self._in(frames, (B, self.net_h, self.net_w, 3), dtype=torch.uint8, name="frames")
            flags |= _lib.CS_FRAME_IN_U8_HWC
        else:
            fr = self._in(frames, (B, 3, self.net_h, self.net_w), name="frames")
        if motion:
            flags |= _lib.CS_FRAME_MOTION
            kt = kc = None
        else:
            kt, kc = self._kp(kp_t, B, "kp_t"), self._kp(kp_can, B, "kp_can")
        if out_u8 is None and out_f32 is None:
            out_u8 = self._new(B, 2 * self.net_h, 2 * self.net_w, 3, dtype=torch.uint8)
        if out_u8 is not None:
            out_u8 = self._in(out_u8, (B, 2 * self.net_h, 2 * self.net_w, 3), dtype=torch.uint8, name="out_u8", output=True)
        if out_f32 is not None:
            out_f32 = self._in(out_f32, (B, 3, 2 * self.net_h, 2 * self.net_w), name="out_f32", output=True)
        self._check(self._lib.cs_frame(self._ctx, fr.data_ptr(), kt.data_ptr() if kt is not None else None,
                                       kc.data_ptr() if kc is not None else None,
                                       out_f32.data_ptr() if out_f32 is not None else None,
                                       out_u8.data_ptr() if out_u8 is not None else None, B, flags, self._stream()))
        return out_u8, out_f32

    # ---- motion extractor M + keypoint transform ------------------------
    HEAD_SLICES = (("kp", 0, 63), ("scale", 63, 64), ("pitch", 64, 130), ("yaw", 130, 196), ("roll", 196, 262), ("t", 262, 265),
                   ("exp", 265, 328))

    def motion(self, img: torch.Tensor) -> torch.Tensor:
        """MotionExtractor.forward (reference motion_extractor.py:33-35): [B,3,net_h,net_w] in [0,1] -> raw heads [B,328]."""
        B = self._batch(img)
        x = self._in(img, (B, 3, self.net_h, self.net_w), name="img")
        heads = self._new(B, _lib.MOTION_HEADS)
        self._check(self._lib.cs_motion(self._ctx, x.data_ptr(), heads.data_ptr(), B, self._stream()))
        return heads

    def motion_dict(self, heads: torch.Tensor):
        """The reference's ret_dct (convnextv2.py:130-141) as views of the heads buffer."""
        return {k: heads[:, a:b] for k, a, b in self.HEAD_SLICES}

    def keypoints(self, heads: torch.Tensor):
        """transform_keypoint (reference can_swap_e2e.py:226-254) -> dict(x_s, x_can = scale * kp, R, deg)."""
        B = int(heads.shape[0])
        h = self._in(heads, (B, _lib.MOTION_HEADS), name="heads")
        x_s, x_can = self._new(B, spec.NUM_KP, 3), self._new(B, spec.NUM_KP, 3)
        R, deg = self._new(B, 3, 3), self._new(B, 3)
        self._check(self._lib.cs_keypoints(self._ctx, h.data_ptr(), x_s.data_ptr(), x_can.data_ptr(), R.data_ptr(), deg.data_ptr(), B,
                                           self._stream()))
        return {"x_s": x_s, "x_can": x_can, "R": R, "deg": deg}

    # ---- paste-back ---------------------------------------------------
    def paste_back(self, img_crop: torch.Tensor, mask_crop: torch.Tensor, M_c2o, img_ori: torch.Tensor, out: Optional[torch.Tensor] = None):
        """prepare_paste_back(if_float=True) + paste_back (reference src/utils/crop.py:515-529), fused and bit-exact with the
        reference's cv2.warpAffine arithmetic.  img_crop [B,hc,wc,3] u8, mask_crop [B,hc,wc] f32, img_ori [B,H,W,3] u8 on the
        device; M_c2o: [B,2..3,3] array-like on the HOST (the reference's float32 crop->original matrices)."""
        import numpy as np
        B, hc, wc = int(img_crop.shape[0]), int(img_crop.shape[1]), int(img_crop.shape[2])
        H, W = int(img_ori.shape[1]), int(img_ori.shape[2])
        if B > _lib.PASTE_MAX_BATCH:                  # the C entry point takes <= 16 matrices per launch: chunk
            if out is None:
                out = self._new(B, H, W, 3, dtype=torch.uint8)
            Mh = np.asarray(M_c2o)
            for lo in range(0, B, _lib.PASTE_MAX_BATCH):
                hi = min(B, lo + _lib.PASTE_MAX_BATCH)
                self.paste_back(img_crop[lo:hi], mask_crop[lo:hi], Mh[lo:hi], img_ori[lo:hi], out=out[lo:hi])
            return out
        crop = self._in(img_crop, (B, hc, wc, 3), dtype=torch.uint8, name="img_crop")
        mask = self._in(mask_crop, (B, hc, wc), name="mask_crop")
        ori = self._in(img_ori, (B, H, W, 3), dtype=torch.uint8, name="img_ori")
        M = np.ascontiguousarray(np.asarray(M_c2o, dtype=np.float64).reshape(B, -1, 3)[:, :2, :]).reshape(B, 6)
        if out is None:
            out = self._new(B, H, W, 3, dtype=torch.uint8)
        out = self._in(out, (B, H, W, 3), dtype=torch.uint8, name="out", output=True)
        self._check(self._lib.cs_paste_back(self._ctx, crop.data_ptr(), mask.data_ptr(), M.ctypes.data_as(C.POINTER(C.c_double)),
                                            ori.data_ptr(), out.data_ptr(), B, hc, wc, H, W, self._stream()))
        return out

    def soft_erosion(self, mask: torch.Tensor, kernel: torch.Tensor, threshold: float = 0.9, iterations: int = 3):
        """SoftErosion.forward (reference src/utils/crop.py:37-47): mask [B,H,W] f32 -> (soft mask [B,H,W] f32, x >= threshold [B,H,W]
        bool); kernel [K,K] f32 = the module's `weight` buffer."""
        B, H, W = (int(v) for v in mask.shape)
        K = int(kernel.shape[-1])
        m = self._in(mask, (B, H, W), name="mask")
        kw = self._in(kernel.reshape(K, K), (K, K), name="kernel")
        out = self._new(B, H, W)
        hard = self._new(B, H, W, dtype=torch.uint8)
        self._check(self._lib.cs_soft_erosion(self._ctx, m.data_ptr(), kw.data_ptr(), out.data_ptr(), hard.data_ptr(), B, H, W, K,
                                              float(threshold), int(iterations), self._stream()))
        return out, hard.bool()

    VALID_PARSE_LABELS = (1, 2, 4, 5, 6, 7, 10, 11, 12)      # reference can_swap_pipeline_e2e.py:48

    def parse_mask(self, logits: torch.Tensor, out_hw=(512, 512), valid_list=VALID_PARSE_LABELS, want_labels: bool = False):
        """The post-processing of the face parser's logits (reference can_swap_pipeline_e2e.py:183-190): bilinear upsample to
        out_hw (align_corners=False) -> argmax over the classes -> isin(valid_list).  logits [B,C,h,w] fp32 on the device ->
        mask [B,H,W] fp32 of 1.0 / 0.0 (what SoftErosion consumes) and, on request, the int32 label map."""
        B, Cc, h, w = (int(v) for v in logits.shape)
        H, W = int(out_hw[0]), int(out_hw[1])
        lg = self._in(logits, (B, Cc, h, w), name="logits")
        valid = 0
        for c in valid_list:
            if 0 <= int(c) < Cc:
                valid |= 1 << int(c)
        mask = self._new(B, H, W)
        labels = self._new(B, H, W, dtype=torch.int32) if want_labels else None
        self._check(self._lib.cs_parse_mask(self._ctx, lg.data_ptr(), B, Cc, h, w, H, W, valid, mask.data_ptr(),
                                            labels.data_ptr() if labels is not None else None, self._stream()))
        return (mask, labels) if want_labels else mask

    def calibrate(self, run_batch):
        """Choose the per-conv activation pre-scales from a representative batch: `run_batch()` must call this engine
        (frame / stage calls) at least once.  Returns the measured max |input| per conv (library order)."""
        self._check(self._lib.cs_calibrate(self._ctx, 1, None, 0))
        try:
            run_batch()
        finally:
            buf = (C.c_float * 1024)()
            self._check(self._lib.cs_calibrate(self._ctx, 0, buf, 1024))
        return [float(v) for v in buf]

    def reset_calibration(self):
        self._check(self._lib.cs_calibrate(self._ctx, 2, None, 0))

    # ---- measurement -------------------------------------------------------------------------------
    PROFILE_FAMILIES = ("conv_tcgen05", "conv_simt", "prep", "stats", "sampling", "other")

    def profile(self, enable: bool):
        self._check(self._lib.cs_profile(self._ctx, 1 if enable else 0))

    def profile_read(self):
        """{family: dict(ms, flops, bytes, launches)} accumulated since profile(True) / the last read."""
        buf = (C.c_double * 24)()
        self._check(self._lib.cs_profile_read(self._ctx, buf))
        return {name: {"ms": buf[4 * i], "flops": buf[4 * i + 1], "bytes": buf[4 * i + 2], "launches": int(buf[4 * i + 3])}
                for i, name in enumerate(self.PROFILE_FAMILIES)}

    def profile_dump(self):
        """Per-launch records since profile(True): list of dict(idx, family, ms, flops, bytes, desc)."""
        buf = C.create_string_buffer(1 << 20)
        self._check(self._lib.cs_profile_dump(self._ctx, buf, len(buf)))
        rows = []
        for line in buf.value.decode().splitlines():
            idx, kind, ms, fl, by, desc = line.split(",", 5)
            rows.append({"idx": int(idx), "family": self.PROFILE_FAMILIES[int(kind)], "ms": float(ms), "flops": float(fl),
                         "bytes": float(by), "desc": desc})
        return rows

    # ---- kernel-level test entry points ---------------------------------------------------------
    def test_conv(self, x_cl, w, bias, pad, act=0, slope=0.0, impl=0):
        """x_cl [B,D,H,W,Cin] channels-last, w [Cout,Cin,KD,KH,KW] -> y [B,Do,Ho,Wo,Cout]."""
        B, D, H, W, Cin = x_cl.shape
        Cout, _, KD, KH, KW = w.shape
        PD, PH, PW = pad
        Do, Ho, Wo = D + 2 * PD - KD + 1, H + 2 * PH - KH + 1, W + 2 * PW - KW + 1
        y = self._new(B, Do, Ho, Wo, Cout)
        x_cl, w = x_cl.contiguous(), w.contiguous()
        b = bias.contiguous() if bias is not None else None
        self._check(self._lib.cs_test_conv(self._ctx, x_cl.data_ptr(), w.data_ptr(), b.data_ptr() if b is not None else None,
                                           y.data_ptr(), B, D, H, W, Cin, Cout, KD, KH, KW, PD, PH, PW, act, float(slope),
                                           impl, self._stream()))
        return y

    def test_conv_up2(self, x_cl, w, bias, act=0, slope=0.0):
        """The conv of nearest-upsample(x, (1,2,2)) in phase form on x itself: x_cl [B,D,H,W,Cin], w [Cout,Cin,KD,3,3] ->
        y [B,D,2H,2W,Cout]."""
        B, D, H, W, Cin = x_cl.shape
        Cout, _, KD, KH, KW = w.shape
        y = self._new(B, D, 2 * H, 2 * W, Cout)
        x_cl, w = x_cl.contiguous(), w.contiguous()
        b = bias.contiguous() if bias is not None else None
        self._check(self._lib.cs_test_conv(self._ctx, x_cl.data_ptr(), w.data_ptr(), b.data_ptr() if b is not None else None,
                                           y.data_ptr(), B, D, H, W, Cin, Cout, KD, KH, KW, KD // 2, 1, 1, act, float(slope), 6,
                                           self._stream()))
        return y

    def test_grid_sample3d(self, inp, grid):
        B, Cc, D, H, W = inp.shape
        out = torch.empty_like(inp)
        self._check(self._lib.cs_test_grid_sample3d(self._ctx, inp.contiguous().data_ptr(), grid.contiguous().data_ptr(),
                                                    out.data_ptr(), B, Cc, D, H, W, self._stream()))
        return out

    def test_instance_stats(self, x, eps=1e-5):
        B, Cc = x.shape[:2]
        S = x[0, 0].numel()
        mean, rstd = self._new(B, Cc), self._new(B, Cc)
        self._check(self._lib.cs_test_instance_stats(self._ctx, x.contiguous().data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                                     B, Cc, S, float(eps), self._stream()))
        return mean, rstd
