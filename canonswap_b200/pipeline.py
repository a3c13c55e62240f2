"""The per-frame loop of `CanSwapPipeline.execute` (reference src/can_swap_pipeline_e2e.py:223-283,
generator part) as a batched, multi-GPU frame pipeline.

Frames are independent given the source identity, so the clip is sharded
round-robin -- frame i -> rank i mod world -- with NO data-path collective; the only message is one
broadcast of the 512-float source identity from rank 0 (`broadcast_identity`).  Each rank streams
its shard through `can_swapper.swap_frames` in batches: pinned host u8 frames -> H2D (copy stream)
-> cs_frame (compute stream) -> D2H of the u8 result (copy stream), double-buffered so copies overlap
the kernels.  Results land in a host array indexed by global frame id.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_indices(n_frames: int, rank: int, world: int) -> List[int]:
    """Global frame ids owned by `rank`: i with i % world == rank (BASELINE config 4)."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world {world}")
    return list(range(rank, n_frames, world))


def batches(indices: Sequence[int], batch: int) -> List[List[int]]:
    if batch < 1:
        raise ValueError("batch must be >= 1")
    return [list(indices[i:i + batch]) for i in range(0, len(indices), batch)]


def broadcast_identity(source_id: Optional[torch.Tensor], device=None, src: int = 0) -> torch.Tensor:
    """One broadcast of the ArcFace identity [1,512] fp32 from `src` (the only collective of the path).
    Without an initialised process group this is the identity function."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        if source_id is None:
            raise ValueError("source_id is required on a single process")
        return source_id if device is None else source_id.to(device)
    if dist.get_rank() == src:
        if source_id is None:
            raise ValueError("source rank must provide source_id")
        buf = source_id.detach().to(torch.float32).reshape(1, 512).clone()
    else:
        buf = torch.empty(1, 512, dtype=torch.float32)
    if device is not None:
        buf = buf.to(device)
    dist.broadcast(buf, src=src)
    return buf


def run_sharded(process_batch: Callable[[List[int]], None], n_frames: int, batch: int, rank: int = 0, world: int = 1) -> int:
    """Drive `process_batch(global_ids)` over this rank's shard; returns the number of frames done."""
    mine = shard_indices(n_frames, rank, world)
    for ids in batches(mine, batch):
        process_batch(ids)
    return len(mine)


class FramePipeline:
    """Host-buffer front end of the hot path: the call a user of the reference pipeline would make.

    swapper: canonswap_b200.modules.can_swapper with weights loaded and the identity set.
    """

    def __init__(self, swapper, net_hw=(256, 256), batch: int = 8):
        self.sw = swapper
        self.batch = batch
        self.net_h, self.net_w = net_hw
        dev = torch.device(swapper.device)
        self.dev = dev
        self.compute = torch.cuda.current_stream(dev)
        self.copy_in = torch.cuda.Stream(dev)
        self.copy_out = torch.cuda.Stream(dev)
        H, W = self.net_h, self.net_w
        self.slots = []
        for _ in range(2):
            self.slots.append({
                "frames": torch.empty(batch, H, W, 3, dtype=torch.uint8, device=dev),
                "x_t": torch.empty(batch, 21, 3, device=dev),
                "x_can": torch.empty(batch, 21, 3, device=dev),
                "out": torch.empty(batch, 2 * H, 2 * W, 3, dtype=torch.uint8, device=dev),
                "in_ready": torch.cuda.Event(), "done": torch.cuda.Event(), "drained": torch.cuda.Event(),
            })
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def run(self, frames_u8: torch.Tensor, x_t: Optional[torch.Tensor], x_can: Optional[torch.Tensor], out_u8: torch.Tensor,
            rank: int = 0, world: int = 1, out_local: bool = False) -> int:
        """frames_u8 [T,H,W,3] u8, x_t/x_can [T,21,3] fp32, out_u8 [T,2H,2W,3] u8 -- all PINNED host tensors.
        x_t = x_can = None: the keypoints are derived on the device by the motion extractor (no motion template on the
        host, reference can_swap_pipeline_e2e.py:101-135,225-243).
        Processes frames i % world == rank; returns how many. Synchronises before returning.
        out_local: out_u8 holds only this rank's frames, in the order of shard_indices (row j = global frame rank + j * world)
        instead of being indexed by the global frame id.
        The kernels run on the stream that is current when run() is called (the ctx is single-stream: do not call run()
        of two pipelines sharing one can_swapper from different streams at the same time)."""
        T = frames_u8.shape[0]
        motion = x_t is None and x_can is None
        mine = shard_indices(T, rank, world)
        contiguous = world == 1
        self.compute = torch.cuda.current_stream(self.dev)          # resolved per call, not at construction
        k = 0
        done = 0
        for ids in batches(mine, self.batch):
            s = self.slots[k % 2]
            k += 1
            b = len(ids)
            with torch.cuda.stream(self.copy_in):
                self.copy_in.wait_event(s["done"])            # previous use of this slot's inputs finished
                if contiguous:
                    lo, hi = ids[0], ids[-1] + 1
                    s["frames"][:b].copy_(frames_u8[lo:hi], non_blocking=True)
                    if not motion:
                        s["x_t"][:b].copy_(x_t[lo:hi], non_blocking=True)
                        s["x_can"][:b].copy_(x_can[lo:hi], non_blocking=True)
                else:
                    for j, i in enumerate(ids):
                        s["frames"][j].copy_(frames_u8[i], non_blocking=True)
                        if not motion:
                            s["x_t"][j].copy_(x_t[i], non_blocking=True)
                            s["x_can"][j].copy_(x_can[i], non_blocking=True)
                s["in_ready"].record(self.copy_in)
            self.h2d_bytes += b * (frames_u8[0].numel() + (0 if motion else 2 * 21 * 3 * 4))
            self.compute.wait_event(s["in_ready"])
            self.compute.wait_event(s["drained"])             # previous D2H of this slot's output finished
            if motion:
                self.sw.swap_frames(s["frames"][:b], out_u8=s["out"][:b])
            else:
                self.sw.swap_frames(s["frames"][:b], s["x_t"][:b], s["x_can"][:b], out_u8=s["out"][:b])
            s["done"].record(self.compute)
            with torch.cuda.stream(self.copy_out):
                self.copy_out.wait_event(s["done"])
                if out_local:
                    out_u8[done:done + b].copy_(s["out"][:b], non_blocking=True)
                elif contiguous:
                    out_u8[ids[0]:ids[-1] + 1].copy_(s["out"][:b], non_blocking=True)
                else:
                    for j, i in enumerate(ids):
                        out_u8[i].copy_(s["out"][j], non_blocking=True)
                s["drained"].record(self.copy_out)
            done += b
            self.d2h_bytes += b * out_u8[0].numel()
        self.copy_out.synchronize()
        self.compute.synchronize()
        return len(mine)


class FullLoopPipeline:
    """The whole per-frame body of `CanSwapPipeline.execute` LOOP C after the front-end (reference
    src/can_swap_pipeline_e2e.py:223-283 with flag_pasteback and flag_do_crop, the defaults):

        x_t, x_can  <- motion extractor + transform_keypoint of the cropped frame     (:112-125, :231-243)
        I_p         <- generator (F, warp, swap, refine, warp_decode) + parse_output  (:242-267)
        mask        <- SoftErosion(21, 0.9, 3)(parsing mask)                          (:42, :275)
        frame       <- paste_back(I_p, M_c2o, full frame, prepare_paste_back(mask))   (:277-282)

    Inputs per frame (what the reference's cropper / face parser hand to the loop): the 256x256 crop, the 512x512 parsing
    mask, the crop->original matrix and the full frame.  Everything stays on the device between the H2D of the inputs and the
    D2H of the pasted frame; the reference moves the keypoints, the mask and the image through numpy per frame.
    """

    def __init__(self, swapper, net_hw=(256, 256), batch: int = 8):
        from .pasteback import SoftErosion
        self.sw = swapper
        self.batch = batch
        self.net_h, self.net_w = net_hw
        self.dev = torch.device(swapper.device)
        self.engine = swapper.engine(net_hw, batch)
        self.soft_mask = SoftErosion(kernel_size=21, threshold=0.9, iterations=3).bind(self.engine)   # pipeline_e2e.py:42
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def run(self, crops_u8: torch.Tensor, parse_masks: torch.Tensor, M_c2o, frames_u8: torch.Tensor, out_u8: torch.Tensor) -> int:
        """crops_u8 [T,h,w,3] u8, parse_masks [T,2h,2w] f32 -- or the face parser's raw logits [T,C,hl,wl] f32, whose
        post-processing (upsample -> argmax -> isin, can_swap_pipeline_e2e.py:183-190) then runs on the device too --,
        frames_u8 / out_u8 [T,H,W,3] u8: PINNED host tensors; M_c2o [T,3,3] (or [T,2,3]) numpy.  Returns the number of frames;
        synchronises before returning."""
        import numpy as np
        T = int(crops_u8.shape[0])
        M = np.asarray(M_c2o)
        st = torch.cuda.current_stream(self.dev)
        for lo in range(0, T, self.batch):
            hi = min(T, lo + self.batch)
            crops = crops_u8[lo:hi].to(self.dev, non_blocking=True)
            masks = parse_masks[lo:hi].to(self.dev, non_blocking=True)
            full = frames_u8[lo:hi].to(self.dev, non_blocking=True)
            self.h2d_bytes += crops.numel() + masks.numel() * 4 + full.numel()
            if masks.dim() == 4:                                                  # logits -> parsing mask on the device
                masks = self.engine.parse_mask(masks, (2 * self.net_h, 2 * self.net_w))
            I_p, _ = self.sw.swap_frames(crops)                                   # keypoints from the motion extractor
            soft, _ = self.soft_mask(masks[:, None])
            pasted = self.engine.paste_back(I_p, soft[:, 0], M[lo:hi], full, out=full)
            out_u8[lo:hi].copy_(pasted, non_blocking=True)
            self.d2h_bytes += pasted.numel()
        st.synchronize()
        return T


class V2IPipeline:
    """The video-to-image variant (reference src/can_swap_pipeline_v2i.py:254-321) on the B200 path: the identity of the
    driving video is swapped onto ONE source image, which is then animated by the driving expressions.

      prepare()    once per source (:86-98 execute_face_canonical + the `i == 0` block :285-304): source -> canonical volume ->
                   swap -> decode -> re-extract motion + appearance of the swapped canonical image.  The loop extracts the
                   appearance volume of that SAME image every frame (:308); it is computed once here and kept resident.
      broadcast()  the one collective: the state (the 32x16xhxw appearance volume, 8.4 MB at 512 px, + ~0.6 KB of keypoints /
                   pose) from rank 0 over NCCL -- no other rank needs the source image, the identity or the swap module.
      run()        frames i % world == rank: x_t_2 = scale * (kp_swap @ R + exp_i) + t (:305) and
                   out = warp_decode(feature, x_swap, x_t_2) (:309) in ONE cs_frame call per batch (CS_FRAME_V2I_FEATURE),
                   driving expressions from pinned host memory in, u8 frames to pinned host memory out.
    """
    STATE_KEYS = ("feature", "x_swap", "kp_swap", "R_swap", "t_swap", "scale_swap")

    def __init__(self, swapper, net_hw=(256, 256), batch: int = 8):
        """net_hw: the size the swapped canonical image is resized to before it is animated -- (256, 256) in the reference
        (can_swap_pipeline_v2i.py:294), whatever the size of the source crop; output frames are twice that."""
        self.sw, self.batch = swapper, batch
        self.net_h, self.net_w = net_hw
        self.dev = torch.device(swapper.device)
        self.state = None
        self.h2d_bytes = self.d2h_bytes = 0

    def prepare(self, source_crop: torch.Tensor, driving_id: torch.Tensor):
        """source_crop [1,3,H,W] fp32 in [0,1] on the device (prepare_source), driving_id [1,512]."""
        sw = self.sw
        x_s_info = sw.get_kp_info(source_crop)                                        # :87
        f_s = sw.extract_feature_3d(source_crop)                                      # :89
        x_s = sw.transform_keypoint(x_s_info)                                         # :90
        x_d = x_s_info["scale"] * x_s_info["kp"]                                      # :94
        f_s_can, occ = sw.warping_module.warp(f_s, x_s, x_d)                          # :97
        f_can_swap = sw.swap_module(f_s_can, driving_id)                              # :286
        swap_can = sw.conv_decode(f_can_swap, occ)                                    # :289
        lr = torch.nn.functional.interpolate(swap_can, size=(self.net_h, self.net_w), mode="bilinear", align_corners=False)   # :294
        x_swap_info = sw.get_kp_info(lr)                                              # :297
        x_swap = sw.transform_keypoint(x_swap_info)                                   # :298
        eng = sw.engine((self.net_h, self.net_w), self.batch)
        R_swap = eng.keypoints(x_s_info["_heads"])["R"]                               # :301 get_rotation_matrix of the SOURCE pose
        t_swap = x_s_info["t"].clone()
        t_swap[..., 2] = 0                                                            # :303
        self.state = {"feature": sw.extract_feature_3d(lr).contiguous(), "x_swap": x_swap.contiguous(),
                      "kp_swap": x_swap_info["kp"].contiguous(), "R_swap": R_swap.contiguous(), "t_swap": t_swap.contiguous(),
                      "scale_swap": x_s_info["scale"].clone().contiguous(), "swap_can": swap_can}
        return self.state

    def _shapes(self):
        h, w = self.net_h // 4, self.net_w // 4
        return {"feature": (1, 32, 16, h, w), "x_swap": (1, 21, 3), "kp_swap": (1, 21, 3), "R_swap": (1, 3, 3), "t_swap": (1, 3),
                "scale_swap": (1, 1)}

    def broadcast(self, src: int = 0):
        """One NCCL broadcast of the per-source state from `src` (a no-op without a process group)."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return self.state
        shapes = self._shapes()
        n = sum(int(torch.tensor(s).prod()) for s in shapes.values())
        dev = self.dev if dist.get_backend() == "nccl" else torch.device("cpu")     # gloo (CPU tests of the host logic): host buffer
        buf = torch.empty(n, dtype=torch.float32, device=dev)
        if dist.get_rank() == src:
            buf.copy_(torch.cat([self.state[k].reshape(-1).float() for k in self.STATE_KEYS]))
        dist.broadcast(buf, src=src)
        st, o = {}, 0
        for k in self.STATE_KEYS:
            m = int(torch.tensor(shapes[k]).prod())
            st[k] = buf[o:o + m].reshape(shapes[k]).clone().to(self.dev if dist.get_backend() == "nccl" else dev)
            o += m
        self.state = st
        return st

    def run(self, exp_driving: torch.Tensor, out_u8: torch.Tensor, rank: int = 0, world: int = 1, out_local: bool = False) -> int:
        """exp_driving [T,21,3] fp32 (x_t_info['exp'] of every driving frame) and out_u8 [T or local,2H,2W,3]: PINNED host
        tensors.  Processes frames i % world == rank; returns how many.  Synchronises before returning."""
        st = self.state
        if st is None:
            raise RuntimeError("V2IPipeline.run before prepare() / broadcast()")
        eng = self.sw.engine((self.net_h, self.net_w), self.batch)
        mine = shard_indices(int(exp_driving.shape[0]), rank, world)
        done = 0
        for ids in batches(mine, self.batch):
            b = len(ids)
            if world == 1:
                delta = exp_driving[ids[0]:ids[-1] + 1].to(self.dev, non_blocking=True)
            else:
                delta = exp_driving[torch.tensor(ids)].to(self.dev, non_blocking=True)
            self.h2d_bytes += delta.numel() * 4
            x_t_2 = st["scale_swap"] * (st["kp_swap"] @ st["R_swap"] + delta) + st["t_swap"]          # :305
            u8, _ = eng.frame(st["feature"], st["x_swap"].expand(b, -1, -1).contiguous(), x_t_2.contiguous(), v2i_feature=True)
            if out_local:
                out_u8[done:done + b].copy_(u8, non_blocking=True)
            elif world == 1:
                out_u8[ids[0]:ids[-1] + 1].copy_(u8, non_blocking=True)
            else:
                for j, i in enumerate(ids):
                    out_u8[i].copy_(u8[j], non_blocking=True)
            self.d2h_bytes += u8.numel()
            done += b
        torch.cuda.current_stream(self.dev).synchronize()
        return len(mine)
