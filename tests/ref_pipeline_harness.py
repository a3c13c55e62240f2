"""Execute the REFERENCE's own per-frame loop text against a drop-in `can_swapper` (test infrastructure).

`CanSwapPipeline.execute` (reference src/can_swap_pipeline_e2e.py:137-350) cannot run headless: it needs ffmpeg, insightface /
onnxruntime, the HF hub, imageio, skimage, matplotlib.  But the part that calls the hot path is plain Python over
`self.can_swapper`:

    make_motion_template   :101-135   get_kp_info -> transform_keypoint -> get_rotation_matrix -> numpy template
    LOOP C                 :223-283   dct2device -> extract_feature_3d -> warp -> conv_decode -> swap_module -> conv_decode ->
                                      refine_module -> warp_decode -> parse_output -> soft_mask -> prepare_paste_back -> paste_back

This module imports the unmodified pipeline file (from /root/reference, or from the oracle/_ref bundle on the GPU box) with
its I/O imports replaced by empty stub modules and `src.can_swap_e2e.can_swapper` replaced by the class under test, calls
`make_motion_template` as written, extracts the LOOP C `for` statement from the source of `execute` with `ast`, and executes
exactly that text.  Nothing of the loop is restated here.
"""
from __future__ import annotations

import ast
import importlib
import os
import sys
import textwrap
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFERENCE = "/root/reference"
BUNDLE = os.path.join(ROOT, "oracle", "_ref")

# modules the pipeline file imports that are absent here or drag in absent packages
_STUBS = ["imageio", "skimage", "skimage.draw", "matplotlib", "matplotlib.pyplot", "insightface_func",
          "insightface_func.face_detect_crop_single", "transformers", "src.utils.cropper", "src.utils.video", "src.utils.io",
          "src.utils.filter", "src.can_swap_e2e"]


def reference_root():
    for r in (REFERENCE, BUNDLE):
        if os.path.isfile(os.path.join(r, "src", "can_swap_pipeline_e2e.py")):
            return r
    return None


class _Stub(types.ModuleType):
    """Any attribute is a do-nothing callable / class, so `from x import a, b, c` succeeds."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (), {"__init__": lambda self, *a, **k: None, "__call__": lambda self, *a, **k: None})


def load_pipeline_module(swapper_cls, name="can_swap_pipeline_e2e"):
    """Import reference src/<name>.py with I/O stubs and `can_swapper` = swapper_cls. Returns (module, source)."""
    root = reference_root()
    if root is None:
        raise FileNotFoundError("neither /root/reference nor the oracle/_ref bundle holds src/can_swap_pipeline_e2e.py")
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "src" or k.startswith("src.")}
    saved.update({k: sys.modules.pop(k) for k in _STUBS if k in sys.modules and not k.startswith("src.")})
    sys.path.insert(0, root)
    try:
        for sname in _STUBS:
            if sname.startswith("src."):
                continue
            sys.modules[sname] = _Stub(sname)
        importlib.import_module("src")                           # the real package, then its stubbed submodules
        importlib.import_module("src.utils")
        for sname in _STUBS:
            if sname.startswith("src."):
                m = _Stub(sname)
                sys.modules[sname] = m
                setattr(sys.modules[sname.rsplit(".", 1)[0]], sname.rsplit(".", 1)[1], m)
        sys.modules["src.can_swap_e2e"].can_swapper = swapper_cls
        mod = importlib.import_module("src." + name)
        src = open(os.path.join(root, "src", name + ".py")).read()
        return mod, src
    finally:
        sys.path.remove(root)
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.") or k in _STUBS]:
            del sys.modules[k]
        sys.modules.update(saved)


def loop_c_source(src: str) -> str:
    """The text of LOOP C: the `for i in track(range(n_frames), ...)` statement inside CanSwapPipeline.execute."""
    tree = ast.parse(src)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "CanSwapPipeline")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "execute")
    loops = [n for n in ast.walk(fn) if isinstance(n, ast.For) and isinstance(n.iter, ast.Call) and
             getattr(n.iter.func, "id", "") == "track" and "extract_feature_3d" in ast.get_source_segment(src, n)]
    assert len(loops) == 1, "LOOP C not found in the reference source"
    seg = ast.get_source_segment(src, loops[0], padded=True)
    return textwrap.dedent(seg)


def run_reference_loop(swapper, soft_mask, frames_u8, source_id, masks, M_c2o_lst, full_frames, pasteback=True):
    """Runs make_motion_template + LOOP C of the reference, as written, with `swapper` as `self.can_swapper`.

    frames_u8: list of [256,256,3] u8 crops; masks: [T,512,512] float tensor on the swapper's device (parsing masks);
    M_c2o_lst: list of 3x3 crop->original matrices; full_frames: list of [H,W,3] u8 frames.
    Returns dict(I_p_lst, I_can_lst, rec_can_lst, I_p_pstbk_lst, template)."""
    import numpy as np
    mod, src = load_pipeline_module(type(swapper))
    loop_src = loop_c_source(src)
    Pipe = mod.CanSwapPipeline
    self = Pipe.__new__(Pipe)                                     # no __init__: that is the I/O the harness replaces
    self.can_swapper = swapper
    self.soft_mask = soft_mask
    mod.track = lambda it, **kw: it                               # rich progress bar -> plain iteration
    n_frames = len(frames_u8)
    I_d_lst = swapper.prepare_videos(frames_u8)                                       # :196
    zeros = [np.zeros((1, 2), np.float32)] * n_frames
    template = Pipe.make_motion_template(self, I_d_lst, zeros, zeros, output_fps=25)   # :197, as written
    inf_cfg = types.SimpleNamespace(flag_pasteback=pasteback, flag_do_crop=pasteback)
    ns = dict(vars(mod))
    ns.update(dict(self=self, n_frames=n_frames, driving_template_dct=template, device=swapper.device, I_d_lst=I_d_lst,
                   source_id=source_id, inf_cfg=inf_cfg, masks=masks, target_M_c2o_lst=M_c2o_lst, driving_rgb_lst=full_frames,
                   I_p_lst=[], I_can_lst=[], rec_can_lst=[], I_p_pstbk_lst=[] if pasteback else None, mask=None, np=np))
    exec(compile(loop_src, "reference:can_swap_pipeline_e2e.py:LOOP_C", "exec"), ns)
    return {k: ns[k] for k in ("I_p_lst", "I_can_lst", "rec_can_lst", "I_p_pstbk_lst")} | {"template": template}


# ---- the video-to-image pipeline (reference src/can_swap_pipeline_v2i.py) ------------------------------------------------
def _fn_node(src, name):
    tree = ast.parse(src)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "CanSwapPipeline")
    return next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == name)


def v2i_source_stage_source(src: str) -> str:
    """The hot-path statements of execute_face_canonical (:86-98): from `I_s = self.can_swapper.prepare_source(...)` up to
    and including `I_p = self.can_swapper.parse_output(out)[0]`, as written."""
    fn = _fn_node(src, "execute_face_canonical")
    segs = [ast.get_source_segment(src, n, padded=True) for n in fn.body]
    first = next(i for i, t in enumerate(segs) if "self.can_swapper.prepare_source" in t)
    last = next(i for i, t in enumerate(segs) if "self.can_swapper.parse_output" in t)
    return textwrap.dedent("\n".join(segs[first:last + 1]))


def v2i_loop_source(src: str) -> str:
    fn = _fn_node(src, "execute")
    loops = [n for n in ast.walk(fn) if isinstance(n, ast.For) and isinstance(n.iter, ast.Call) and
             getattr(n.iter.func, "id", "") == "track" and "warp_decode" in ast.get_source_segment(src, n)]
    assert len(loops) == 1, "the v2i frame loop was not found in the reference source"
    return textwrap.dedent(ast.get_source_segment(src, loops[0], padded=True))


def run_reference_v2i(swapper, source_crop_u8, driving_id, driving_crops_u8, source_original, source_M_c2o, mask_ori_float):
    """execute_face_canonical's hot statements, make_motion_template and the frame loop of the reference v2i pipeline, as
    written, with `swapper` as self.can_swapper.  Returns dict(I_p_lst, I_can_lst, I_p_pstbk_lst, template, x_s_info, f_s_can)."""
    import numpy as np
    import torch
    mod, src = load_pipeline_module(type(swapper), "can_swap_pipeline_v2i")
    Pipe = mod.CanSwapPipeline
    self = Pipe.__new__(Pipe)
    self.can_swapper = swapper
    mod.track = lambda it, **kw: it
    ns = dict(vars(mod))
    ns.update(dict(self=self, img_crop_256x256=source_crop_u8, np=np, torch=torch))
    exec(compile(v2i_source_stage_source(src), "reference:can_swap_pipeline_v2i.py:execute_face_canonical", "exec"), ns)
    n_frames = len(driving_crops_u8)
    I_d_lst = swapper.prepare_videos(driving_crops_u8)
    zeros = [np.zeros((1, 2), np.float32)] * n_frames
    template = Pipe.make_motion_template(self, I_d_lst, zeros, zeros, output_fps=25)
    ns.update(dict(n_frames=n_frames, driving_template_dct=template, device=swapper.device, driving_id=driving_id,
                   occ_map=ns["occ_map"], f_s_can=ns["f_s_can"], x_s_info=ns["x_s_info"], source_original=source_original,
                   source_M_c2o=source_M_c2o, mask_ori_float=mask_ori_float, I_can_lst=[], I_p_lst=[], I_p_pstbk_lst=[]))
    exec(compile(v2i_loop_source(src), "reference:can_swap_pipeline_v2i.py:frame_loop", "exec"), ns)
    return {k: ns[k] for k in ("I_p_lst", "I_can_lst", "I_p_pstbk_lst", "x_s_info", "f_s_can")} | {"template": template}
