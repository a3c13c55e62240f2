"""Motion extractor M + keypoint transform.

CPU: the oracle restatement (oracle/canonswap_oracle.py: motion_extractor, transform_keypoint, motion_keypoints) against
the golden fixtures written by the UNMODIFIED reference module (tests/golden/make_golden.py motion) and, when
/root/reference is present, against the live reference module and its state_dict layout.
GPU: the CUDA path (cs_motion / cs_keypoints / cs_frame with CS_FRAME_MOTION) against the oracle and the goldens.

Tolerance: the heads feed keypoints in [-1, 1] normalised coordinates; measured on B200 they are within 2e-6 (kp) /
1e-5 (angle logits) of the oracle, the tests ask for 1e-4 * max(1, range) -- ten times tighter than the hot path's 1e-3.
The generator itself is very sensitive to its keypoints on this fixture (a 1e-5 shift of x_t moves the image by ~1e-2 at
the mask boundaries, tools/motion_err.py), so the fused call is checked in two parts: keypoints against the oracle, and
the image against the oracle generator fed the SAME (device-derived) keypoints at the hot path's 1e-3.
"""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import has_reference
from canonswap_b200 import spec, synth
from oracle import canonswap_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-3
HEAD_TOL = 1e-4
KEYS = ("pitch", "yaw", "roll", "t", "exp", "scale", "kp")


@pytest.fixture(scope="module")
def motion_w():
    return synth.synth_motion_state_dict()


def _golden(tag):
    return {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLDEN, f"motion_{tag}.npz")).items()}


@pytest.mark.parametrize("tag,T,hw", [("b2_256", 2, 256), ("b1_128", 1, 128)])
def test_oracle_motion_matches_reference_golden(motion_w, tag, T, hw):
    g = _golden(tag)
    inp = synth.synth_inputs(T, hw)
    mk = O.motion_keypoints(motion_w, inp["frames"])
    for k in KEYS:
        assert (mk["info"][k] - g[k]).abs().max().item() <= 1e-5 * max(1.0, g[k].abs().max().item()), k
    assert (mk["x_t"] - g["x_s"]).abs().max().item() <= 1e-5
    assert (mk["x_can"] - g["x_can"]).abs().max().item() <= 1e-5
    assert (mk["R"] - g["R"]).abs().max().item() <= 1e-6
    assert (mk["deg"] - g["deg"]).abs().max().item() <= 1e-4


def test_motion_spec_is_a_complete_state_dict(motion_w):
    sp = spec.motion_extractor_spec()
    assert list(sp.keys()) == list(motion_w.keys())
    assert all(tuple(motion_w[k].shape) == tuple(s) for k, s in sp.items())
    assert sum(n for _, n in spec.MOTION_HEADS) == 328


@pytest.mark.reference
@pytest.mark.skipif(not has_reference(), reason="/root/reference not present")
def test_oracle_motion_matches_live_reference(motion_w):
    sys.path.insert(0, GOLDEN)
    import make_golden
    m = make_golden.build_reference_motion()
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == [(k, tuple(s)) for k, s in spec.motion_extractor_spec().items()]
    m.load_state_dict(motion_w, strict=True)
    inp = synth.synth_inputs(1, 256, seed=77)
    ref = make_golden.reference_motion(m, inp["frames"])
    mk = O.motion_keypoints(motion_w, inp["frames"])
    for k in KEYS:
        assert (mk["info"][k] - ref[k]).abs().max().item() <= 1e-5 * max(1.0, ref[k].abs().max().item()), k
    assert (mk["x_t"] - ref["x_s"]).abs().max().item() <= 1e-5
    assert (mk["x_can"] - ref["x_can"]).abs().max().item() <= 1e-5


# ---- GPU ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def eng256(synth_w, motion_w):
    from canonswap_b200.engine import Engine
    w = dict(synth_w)
    w[spec.MOTION_NET] = motion_w
    eng = Engine(w, net_hw=(256, 256), max_batch=2, device=0)
    yield eng
    eng.close()


@pytest.mark.gpu
def test_motion_heads_match_oracle_and_golden(eng256, motion_w):
    g = _golden("b2_256")
    inp = synth.synth_inputs(2, 256)
    heads = eng256.motion(inp["frames"].cuda())
    d = eng256.motion_dict(heads)
    ref = O.motion_extractor(motion_w, inp["frames"])
    for k in KEYS:
        got = d[k].cpu()
        for name, want in (("oracle", ref[k]), ("golden", g[k])):
            err = (got - want).abs().max().item()
            assert err <= HEAD_TOL * max(1.0, want.abs().max().item()), (k, name, err)


@pytest.mark.gpu
def test_keypoint_transform_matches_oracle(eng256, motion_w):
    """cs_keypoints alone: fed with the ORACLE's heads so that only the transform is compared."""
    inp = synth.synth_inputs(2, 256, seed=5)
    info = O.motion_extractor(motion_w, inp["frames"])
    heads = torch.cat([info[k] for k in ("kp", "scale", "pitch", "yaw", "roll", "t", "exp")], dim=1)
    assert heads.shape == (2, 328)
    want = O.motion_keypoints(motion_w, inp["frames"])
    got = eng256.keypoints(heads.cuda())
    assert (got["x_s"].cpu() - want["x_t"]).abs().max().item() <= 1e-5
    assert (got["x_can"].cpu() - want["x_can"]).abs().max().item() <= 1e-6
    assert (got["R"].cpu() - want["R"]).abs().max().item() <= 1e-6
    assert (got["deg"].cpu() - want["deg"]).abs().max().item() <= 1e-3


@pytest.mark.gpu
def test_motion_extractor_mirror_module_and_kp_info(synth_w, motion_w):
    """The reference-facing surface: can_swapper.motion_extractor(x) / get_kp_info / transform_keypoint."""
    from canonswap_b200.modules import can_swapper
    w = dict(synth_w)
    w[spec.MOTION_NET] = motion_w
    sw = can_swapper(weights=w, device_id=0, max_batch=2)
    inp = synth.synth_inputs(2, 256)
    x = inp["frames"].cuda()
    want = O.motion_keypoints(motion_w, inp["frames"])
    info = sw.get_kp_info(x)
    assert info["kp"].shape == (2, 21, 3) and info["exp"].shape == (2, 21, 3) and info["pitch"].shape == (2, 1)
    assert (info["pitch"].cpu()[:, 0] - want["deg"][:, 0]).abs().max().item() <= 5e-2      # degrees
    x_s = sw.transform_keypoint(info)
    assert (x_s.cpu() - want["x_t"]).abs().max().item() <= HEAD_TOL


@pytest.mark.gpu
def test_frame_with_motion_matches_oracle(synth_w, motion_w):
    """cs_frame with CS_FRAME_MOTION: keypoints derived on the device from the frames.  Keypoints within 1e-4 of the
    oracle's; image within 1e-3 of the oracle generator fed the same keypoints; and the fused call gives the same result as
    cs_motion -> cs_keypoints -> cs_frame (up to the order of the GRN atomics)."""
    from canonswap_b200.engine import Engine
    w = dict(synth_w)
    w[spec.MOTION_NET] = motion_w
    inp = synth.synth_inputs(2, 128)
    mk = O.motion_keypoints(motion_w, inp["frames"])
    eng = Engine(w, net_hw=(128, 128), max_batch=2, device=0)
    try:
        eng.set_identity(inp["source_id"].cuda())
        kp = eng.keypoints(eng.motion(inp["frames"].cuda()))
        assert (kp["x_s"].cpu() - mk["x_t"]).abs().max().item() <= HEAD_TOL
        assert (kp["x_can"].cpu() - mk["x_can"]).abs().max().item() <= HEAD_TOL
        ref = O.frame(synth_w, inp["frames"], kp["x_s"].cpu(), kp["x_can"].cpu(), inp["source_id"])["out"]
        out = torch.empty(2, 3, 256, 256, device="cuda")
        eng.frame(inp["frames"].cuda(), out_f32=out, motion=True)
        d = (out.cpu() - ref).abs().max().item()
        assert d <= TOL, d
        o1 = torch.empty_like(out)
        eng.frame(inp["frames"].cuda(), kp["x_s"], kp["x_can"], out_f32=o1)
        assert (o1 - out).abs().max().item() <= TOL      # same kernels; GRN sums use fp64 atomics (order may vary)
        # graph replay (two lanes) gives the same bytes as the eager call
        from canonswap_b200 import _lib
        eng.set_option(_lib.CS_OPT_USE_GRAPH, 1)
        o2 = torch.empty_like(out)
        for _ in range(3):
            eng.frame(inp["frames"].cuda(), out_f32=o2, motion=True)
        assert torch.equal(o2, out)
    finally:
        eng.close()


@pytest.mark.gpu
def test_motion_without_weights_fails_loudly(synth_w):
    from canonswap_b200.engine import Engine, CanonSwapError
    eng = Engine(synth_w, net_hw=(128, 128), max_batch=1, device=0)
    try:
        with pytest.raises(CanonSwapError):
            eng.motion(torch.zeros(1, 3, 128, 128, device="cuda"))
    finally:
        eng.close()
