"""Loader for oracle/_ref (see make_ref.py): the UNMODIFIED reference modules composed exactly as the per-frame loop
composes them (reference src/can_swap_pipeline_e2e.py:242-263, src/can_swap_e2e.py:60-68,286-312).

TEST / MEASUREMENT INFRASTRUCTURE ONLY -- imported by bench.py's reference arm / cpu_baseline leg and by tests/.
"""
from __future__ import annotations

import importlib
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(DST, "src", "modules", "warping_network.py"))


def _import(name):
    """Import `src.modules.X` from the bundle without leaving the bundle on sys.path (and without picking up another `src`)."""
    saved = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, DST)
    try:
        return importlib.import_module(name)
    finally:
        sys.path.remove(DST)
        bundle = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
        for k in bundle:
            del sys.modules[k]
        sys.modules.update(saved)
        _CACHE.update(bundle)


_CACHE = {}


def build_modules(weights, device="cpu"):
    """The five hot-path networks of can_swapper.__init__ (can_swap_e2e.py:60-68) with `weights` (combined_weights.pth layout)
    loaded strictly, in eval mode."""
    import yaml
    if not available():
        raise FileNotFoundError("oracle/_ref is missing: run `python oracle/make_ref.py` where /root/reference exists")
    sys.modules.update(_CACHE)
    try:
        afe = _import("src.modules.appearance_feature_extractor")
        sys.modules.update(_CACHE)
        wn = _import("src.modules.warping_network")
        sys.modules.update(_CACHE)
        sg = _import("src.modules.spade_generator")
        sys.modules.update(_CACHE)
        am = _import("src.modules.adaptive_modulate")
    finally:
        for k in list(sys.modules):
            if (k == "src" or k.startswith("src.")) and k in _CACHE:
                del sys.modules[k]
    cfg = yaml.safe_load(open(os.path.join(DST, "src", "config", "models.yaml")))["model_params"]
    cfg["spade_generator_params"]["upscale"] = 2                       # can_swap_e2e.py:62
    mods = {
        "appearance_feature_extractor": afe.AppearanceFeatureExtractor(**cfg["appearance_feature_extractor_params"]),
        "warping_module": wn.WarpingNetwork(**cfg["warping_module_params"]),
        "spade_generator": sg.SPADEDecoder(**cfg["spade_generator_params"]),
        "transfer": am.transfer_model2(),
        "refine": am.G3d(),
    }
    for name, m in mods.items():
        m.load_state_dict(weights[name], strict=True)
        m.to(device).eval()
    return mods


def frame(mods, I_s, x_t, x_can, source_id):
    """LOOP C core path with the reference modules (pipeline_e2e.py:242-263, debug decodes off) -> image [B,3,2H,2W]."""
    F_, W_, G_, T_, R_ = (mods[k] for k in ("appearance_feature_extractor", "warping_module", "spade_generator", "transfer",
                                            "refine"))
    with torch.no_grad():
        f_s = F_(I_s)                                                          # :242
        f_can, _occ = W_.warp(f_s, x_t, x_can)                                 # :244
        f_swap = T_(f_can, source_id.expand(I_s.shape[0], -1))                 # :253
        f_swap = R_(f_swap)                                                    # :262
        ret = W_(f_swap, kp_source=x_can, kp_driving=x_t)                      # :263 -> can_swap_e2e.py:298
        return G_(feature=ret["out"])                                          # can_swap_e2e.py:300
