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


_MODS = {}


def _load_bundle():
    """Import the bundle's `src.modules.*` once, without leaving `oracle/_ref` on sys.path or its `src` package in sys.modules
    (another `src` -- e.g. /root/reference itself in the authoring container -- must stay importable)."""
    if _MODS:
        return _MODS
    if not available():
        raise FileNotFoundError("oracle/_ref is missing: run `python oracle/make_ref.py` where /root/reference exists")
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "src" or k.startswith("src.")}
    sys.path.insert(0, DST)
    try:
        for name in ("appearance_feature_extractor", "warping_network", "spade_generator", "adaptive_modulate", "motion_extractor"):
            _MODS[name] = importlib.import_module("src.modules." + name)
    finally:
        sys.path.remove(DST)
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    return _MODS


def build_modules(weights, device="cpu"):
    """The five hot-path networks of can_swapper.__init__ (can_swap_e2e.py:60-68) with `weights` (combined_weights.pth layout)
    loaded strictly, in eval mode."""
    import yaml
    m = _load_bundle()
    cfg = yaml.safe_load(open(os.path.join(DST, "src", "config", "models.yaml")))["model_params"]
    cfg["spade_generator_params"]["upscale"] = 2                       # can_swap_e2e.py:62
    mods = {
        "appearance_feature_extractor": m["appearance_feature_extractor"].AppearanceFeatureExtractor(**cfg["appearance_feature_extractor_params"]),
        "warping_module": m["warping_network"].WarpingNetwork(**cfg["warping_module_params"]),
        "spade_generator": m["spade_generator"].SPADEDecoder(**cfg["spade_generator_params"]),
        "transfer": m["adaptive_modulate"].transfer_model2(),
        "refine": m["adaptive_modulate"].G3d(),
    }
    for name, mod in mods.items():
        mod.load_state_dict(weights[name], strict=True)
        mod.to(device).eval()
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
