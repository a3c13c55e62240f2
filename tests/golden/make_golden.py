"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules from /root/reference.

Run in the authoring container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py

The reference `src/modules/*` classes are instantiated from the reference's own
`src/config/models.yaml`, loaded (strict) with the seeded synthetic state_dicts of
`canonswap_b200.synth`, and composed exactly as the per-frame loop does
(reference can_swap_pipeline_e2e.py:242-263, can_swap_e2e.py:286-312).  For every stage a
strided sample of the output (<= 4096 values) plus its mean/abs-mean is stored, so the fixtures
stay small; the final image is stored at full size for the smallest case.
"""
import os
import sys

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from canonswap_b200 import synth  # noqa: E402

STAGES = ["f_s", "f_can", "occ_can", "f_swap", "f_refine", "occ", "deformation", "warp_out", "out"]


def build_reference_modules():
    sys.path.insert(0, REF)
    from src.modules.appearance_feature_extractor import AppearanceFeatureExtractor
    from src.modules.warping_network import WarpingNetwork
    from src.modules.spade_generator import SPADEDecoder
    from src.modules.adaptive_modulate import transfer_model2, G3d
    cfg = yaml.safe_load(open(os.path.join(REF, "src/config/models.yaml")))["model_params"]
    cfg["spade_generator_params"]["upscale"] = 2                       # can_swap_e2e.py:62
    mods = {
        "appearance_feature_extractor": AppearanceFeatureExtractor(**cfg["appearance_feature_extractor_params"]),
        "warping_module": WarpingNetwork(**cfg["warping_module_params"]),
        "spade_generator": SPADEDecoder(**cfg["spade_generator_params"]),
        "transfer": transfer_model2(),
        "refine": G3d(),
    }
    for m in mods.values():
        m.eval()
    return mods


def reference_frame(mods, I_s, x_t, x_can, source_id):
    """LOOP C with the reference modules (pipeline_e2e.py:242-263)."""
    F_, W_, G_, T_, R_ = (mods[k] for k in ("appearance_feature_extractor", "warping_module",
                                            "spade_generator", "transfer", "refine"))
    r = {}
    with torch.no_grad():
        r["f_s"] = F_(I_s)
        r["f_can"], r["occ_can"] = W_.warp(r["f_s"], x_t, x_can)
        r["f_swap"] = T_(r["f_can"], source_id.expand(I_s.shape[0], -1))
        r["f_refine"] = R_(r["f_swap"])
        ret = W_(r["f_refine"], kp_source=x_can, kp_driving=x_t)
        r["occ"], r["deformation"], r["warp_out"] = ret["occlusion_map"], ret["deformation"], ret["out"]
        r["out"] = G_(feature=ret["out"])
    return r


def build_reference_motion():
    """The reference MotionExtractor (src/modules/motion_extractor.py) with models.yaml's parameters."""
    sys.path.insert(0, REF)
    from src.modules.motion_extractor import MotionExtractor
    cfg = yaml.safe_load(open(os.path.join(REF, "src/config/models.yaml")))["model_params"]
    return MotionExtractor(**cfg["motion_extractor_params"]).eval()


def reference_motion(m, I):
    """get_kp_info + transform_keypoint + x_can with the reference's own functions (can_swap_e2e.py:174-254,
    camera.py:14-73, pipeline_e2e.py:112-125,242)."""
    sys.path.insert(0, REF)
    from src.utils.camera import get_rotation_matrix, headpose_pred_to_degree
    with torch.no_grad():
        info = m(I)
    bs = I.shape[0]
    pitch, yaw, roll = (headpose_pred_to_degree(info[k]) for k in ("pitch", "yaw", "roll"))
    rot = get_rotation_matrix(pitch, yaw, roll)
    x_s = info["kp"].view(bs, 21, 3) @ rot + info["exp"].view(bs, 21, 3)
    x_s = x_s * info["scale"][..., None]
    x_s[:, :, 0:2] += info["t"][:, None, 0:2]
    x_can = info["scale"][..., None] * info["kp"].view(bs, 21, 3)
    r = dict(info)
    r.update(x_s=x_s, x_can=x_can, R=rot, deg=torch.stack([pitch, yaw, roll], 1))
    return r


def sample(t, n=4096):
    flat = t.reshape(-1)
    step = max(1, flat.numel() // n)
    return flat[::step][:n].contiguous()


def main_motion():
    """motion extractor: full outputs of the reference module + keypoint transform"""
    mm = build_reference_motion()
    mm.load_state_dict(synth.synth_motion_state_dict(), strict=True)
    for tag, T, hw in (("b2_256", 2, 256), ("b1_128", 1, 128)):
        inp = synth.synth_inputs(T, hw)
        r = reference_motion(mm, inp["frames"])
        np.savez_compressed(os.path.join(HERE, f"motion_{tag}.npz"), **{k: v.numpy().astype(np.float32) for k, v in r.items()})
        print("motion", tag, {k: float(v.abs().mean()) for k, v in r.items()})


def pasteback_case(seed, hc=96, wc=96, H=160, W=200):
    """Seeded inputs of one paste-back frame: crop image, soft mask, crop->original matrix (float32 3x3), full frame."""
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (hc, wc, 3), dtype=np.uint8)
    ori = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:hc, 0:wc].astype(np.float32)
    mask = np.clip(1.4 - np.sqrt(((yy - hc / 2) / (0.4 * hc)) ** 2 + ((xx - wc / 2) / (0.35 * wc)) ** 2), 0, 1).astype(np.float32)
    ang, sc = rng.uniform(-0.6, 0.6), rng.uniform(0.5, 1.8)
    M = np.array([[sc * np.cos(ang), -sc * np.sin(ang), rng.uniform(-30, 90)],
                  [sc * np.sin(ang), sc * np.cos(ang), rng.uniform(-30, 60)], [0, 0, 1]], dtype=np.float32)
    return img, mask, M, ori


def main_pasteback():
    """paste-back: the reference's own prepare_paste_back / paste_back (src/utils/crop.py:515-529)"""
    sys.path.insert(0, REF)
    from src.utils.crop import prepare_paste_back, paste_back
    for seed in (0, 1, 2):
        img, mask, M, ori = pasteback_case(seed)
        m3 = np.stack([mask] * 3, axis=-1)
        mask_ori = prepare_paste_back(m3, M, dsize=(ori.shape[1], ori.shape[0]), if_float=True)
        out = paste_back(img, M, ori, mask_ori)
        np.savez_compressed(os.path.join(HERE, f"pasteback_{seed}.npz"), out=out, mask_ori=mask_ori[..., 0].astype(np.float32))
        print("pasteback", seed, out.shape, float(mask_ori.mean()))


def main():
    mods = build_reference_modules()
    W = synth.synth_weights()
    for name, m in mods.items():
        m.load_state_dict(W[name], strict=True)
    for tag, T, hw in (("b1_128", 1, 128), ("b2_128", 2, 128), ("b1_256", 1, 256)):
        inp = synth.synth_inputs(T, hw)
        r = reference_frame(mods, inp["frames"], inp["x_t"], inp["x_can"], inp["source_id"])
        out = {}
        for k in STAGES:
            out[k + "_sample"] = sample(r[k]).numpy()
            out[k + "_mean"] = np.float64(r[k].double().mean().item())
            out[k + "_absmean"] = np.float64(r[k].double().abs().mean().item())
            out[k + "_shape"] = np.array(r[k].shape)
        if tag == "b1_128":
            out["out_full"] = r["out"].numpy().astype(np.float32)
        np.savez_compressed(os.path.join(HERE, f"frame_{tag}.npz"), **out)
        print(tag, {k: float(out[k + "_absmean"]) for k in STAGES})


if __name__ == "__main__":
    if "pasteback" in sys.argv[1:]:
        main_pasteback()
    elif "motion" in sys.argv[1:]:
        main_motion()
    else:
        main()
        main_motion()
        main_pasteback()
