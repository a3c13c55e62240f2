"""Paste-back of the swapped crop into the full frame: reference src/utils/crop.py:515-529,
cv2.warpAffine INTER_LINEAR (src/utils/crop.py:49-63), called per frame at src/can_swap_pipeline_e2e.py:277-282.

Byte work: the bar is BIT-EXACT.
CPU: the numpy oracle (oracle/pasteback_oracle.py) against cv2 itself (present in this image), against the golden outputs of
the reference's own functions (tests/golden/pasteback_*.npz, make_golden.py pasteback) and, when /root/reference is
present, against those functions live.  GPU: the fused CUDA kernel (cs_paste_back) against the oracle, bit for bit, over
random similarity transforms, crops hanging over the frame edge, identity / degenerate matrices and a 1080p frame.
"""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import has_reference
from oracle import pasteback_oracle as P

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLDEN)
from make_golden import pasteback_case  # noqa: E402


def _rand_case(seed, hc, wc, H, W, kind="similarity"):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (hc, wc, 3), dtype=np.uint8)
    ori = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    mask = rng.random((hc, wc), dtype=np.float32)
    mask[mask > 0.7] = 1.0
    mask[mask < 0.1] = 0.0
    if kind == "identity":
        M = np.eye(3, dtype=np.float32)
    elif kind == "offscreen":
        M = np.array([[1.0, 0, W + 50], [0, 1.0, 10], [0, 0, 1]], dtype=np.float32)
    elif kind == "singular":
        M = np.array([[1.0, 2.0, 3.0], [2.0, 4.0, 1.0], [0, 0, 1]], dtype=np.float32)
    elif kind == "edge":
        M = np.array([[1.3, 0.2, -0.4 * wc], [-0.15, 1.25, H - 0.5 * hc], [0, 0, 1]], dtype=np.float32)
    else:
        ang, sc = rng.uniform(-3.1, 3.1), rng.uniform(0.3, 2.5)
        M = np.array([[sc * np.cos(ang), -sc * np.sin(ang), rng.uniform(-0.3 * W, 0.9 * W)],
                      [sc * np.sin(ang), sc * np.cos(ang), rng.uniform(-0.3 * H, 0.9 * H)], [0, 0, 1]], dtype=np.float32)
    return img, mask, M, ori


def test_oracle_warp_is_bit_exact_with_cv2():
    cv2 = pytest.importorskip("cv2")
    cv2.setNumThreads(0)
    for seed in range(6):
        img, mask, M, ori = _rand_case(seed, 64, 80, 120, 150)
        dsize = (ori.shape[1], ori.shape[0])
        assert np.array_equal(P.warp_affine_u8(img, M, dsize), cv2.warpAffine(img, M[:2, :], dsize, flags=cv2.INTER_LINEAR))
        m3 = np.stack([mask] * 3, -1)
        assert np.array_equal(P.warp_affine_f32(m3, M, dsize), cv2.warpAffine(m3, M[:2, :], dsize, flags=cv2.INTER_LINEAR))
    for kind in ("identity", "offscreen", "edge"):
        img, mask, M, ori = _rand_case(9, 48, 40, 90, 70, kind)
        dsize = (70, 90)
        assert np.array_equal(P.warp_affine_u8(img, M, dsize), cv2.warpAffine(img, M[:2, :], dsize, flags=cv2.INTER_LINEAR)), kind


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_oracle_matches_reference_golden(seed):
    g = np.load(os.path.join(GOLDEN, f"pasteback_{seed}.npz"))
    img, mask, M, ori = pasteback_case(seed)
    m3 = np.stack([mask] * 3, -1)
    mask_ori = P.prepare_paste_back(m3, M, (ori.shape[1], ori.shape[0]), if_float=True)
    assert np.array_equal(mask_ori[..., 0], g["mask_ori"])
    assert np.array_equal(P.paste_back(img, M, ori, mask_ori), g["out"])
    assert np.array_equal(P.paste_back_frame(img, mask, M, ori), g["out"])


@pytest.mark.reference
@pytest.mark.skipif(not has_reference(), reason="/root/reference not present")
def test_oracle_matches_live_reference_functions():
    sys.path.insert(0, "/root/reference")
    from src.utils.crop import prepare_paste_back, paste_back
    for seed in range(4):
        img, mask, M, ori = _rand_case(100 + seed, 96, 96, 180, 240)
        m3 = np.stack([mask] * 3, -1)
        mo = prepare_paste_back(m3, M, dsize=(ori.shape[1], ori.shape[0]), if_float=True)
        assert np.array_equal(P.paste_back_frame(img, mask, M, ori), paste_back(img, M, ori, mo))


# ---- GPU -----------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def eng():
    from canonswap_b200.engine import Engine
    e = Engine(None, net_hw=(128, 128), max_batch=1, device=0)
    yield e
    e.close()


def _gpu(eng, cases):
    crop = torch.from_numpy(np.stack([c[0] for c in cases])).cuda()
    mask = torch.from_numpy(np.stack([c[1] for c in cases])).cuda()
    ori = torch.from_numpy(np.stack([c[3] for c in cases])).cuda()
    M = np.stack([c[2] for c in cases])
    return eng.paste_back(crop, mask, M, ori).cpu().numpy()


@pytest.mark.gpu
def test_paste_back_kernel_is_bit_exact_with_oracle(eng):
    cases = [_rand_case(s, 96, 112, 200, 260) for s in range(8)]
    out = _gpu(eng, cases)
    for i, c in enumerate(cases):
        assert np.array_equal(out[i], P.paste_back_frame(*c)), i


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["identity", "offscreen", "edge", "singular"])
def test_paste_back_kernel_edge_cases(eng, kind):
    c = _rand_case(5, 64, 48, 100, 140, kind)
    assert np.array_equal(_gpu(eng, [c])[0], P.paste_back_frame(*c))


@pytest.mark.gpu
def test_paste_back_more_frames_than_one_launch_takes(eng):
    cases = [_rand_case(200 + s, 40, 40, 64, 80) for s in range(19)]          # > CS_PASTE_MAX_BATCH = 16: chunked by the binding
    out = _gpu(eng, cases)
    for i, c in enumerate(cases):
        assert np.array_equal(out[i], P.paste_back_frame(*c)), i


@pytest.mark.gpu
def test_paste_back_kernel_odd_width_scalar_path(eng):
    """W % 4 != 0 takes the one-pixel-per-thread kernel (the 4-pixel kernel needs word-aligned rows)."""
    cases = [_rand_case(s, 50, 70, 97, 141) for s in (11, 12)]
    out = _gpu(eng, cases)
    for i, c in enumerate(cases):
        assert np.array_equal(out[i], P.paste_back_frame(*c)), i


@pytest.mark.gpu
def test_paste_back_kernel_matches_reference_golden(eng):
    for seed in (0, 1, 2):
        g = np.load(os.path.join(GOLDEN, f"pasteback_{seed}.npz"))
        assert np.array_equal(_gpu(eng, [pasteback_case(seed)])[0], g["out"])


@pytest.mark.gpu
def test_paste_back_1080p_frame_and_in_place(eng):
    """The reference's operating point: a 512x512 crop pasted into a 1920x1080 frame; `out` may alias `img_ori`."""
    c = _rand_case(42, 512, 512, 1080, 1920)
    want = P.paste_back_frame(*c)
    assert np.array_equal(_gpu(eng, [c])[0], want)
    crop, mask = torch.from_numpy(c[0][None]).cuda(), torch.from_numpy(c[1][None]).cuda()
    ori = torch.from_numpy(c[3][None]).cuda()
    eng.paste_back(crop, mask, c[2][None], ori, out=ori)
    assert np.array_equal(ori.cpu().numpy()[0], want)


def _face_mask(H=256, W=256, seed=0):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    m = (((yy - H * 0.52) / (0.33 * H)) ** 2 + ((xx - W * 0.48) / (0.27 * W)) ** 2 < 1).astype(np.float32)
    m[rng.random((H, W)) < 0.002] = 0                      # pin-holes, as a parsing mask has
    return torch.from_numpy(m)[None, None]


@pytest.mark.reference
@pytest.mark.skipif(not has_reference(), reason="/root/reference not present")
def test_soft_erosion_oracle_matches_reference_module():
    sys.path.insert(0, "/root/reference")
    from src.utils.crop import SoftErosion as RefSE
    from canonswap_b200.pasteback import SoftErosion
    x = _face_mask()
    ref = RefSE(kernel_size=21, threshold=0.9, iterations=3)
    want, wmask = ref(x.clone())
    got, gmask, _ = P.soft_erosion(x, 21, 0.9, 3)
    assert torch.equal(got, want) and torch.equal(gmask, wmask)
    assert torch.equal(SoftErosion(21, 0.9, 3).weight, ref.weight)     # the mirror builds the same buffer


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [(21, 0.9, 3), (21, 0.9, 2), (15, 0.6, 1)])
def test_soft_erosion_kernel_matches_oracle(eng, cfg):
    """Float work (a 21x21 convolution in a different summation order): within 1e-5 wherever the pre-threshold value is not
    within 1e-5 of the threshold, where the hard decision may flip."""
    from canonswap_b200.pasteback import SoftErosion
    K, thr, it = cfg
    x = _face_mask(seed=3)
    want, wmask, pre = P.soft_erosion(x, K, thr, it)
    se = SoftErosion(K, thr, it).bind(eng)
    got, gmask = se(x.cuda())
    safe = (pre - thr).abs() > 1e-5
    assert safe.float().mean().item() > 0.99
    assert ((got.cpu() - want).abs()[safe]).max().item() <= 1e-5
    assert torch.equal(gmask.cpu()[safe], wmask[safe])


@pytest.mark.gpu
def test_full_loop_pipeline_against_oracle_chain(synth_w):
    """FullLoopPipeline = motion extractor -> generator -> parse_output -> SoftErosion -> paste-back, host buffers in and out.
    Checked stage-wise against the oracle chain fed the device's own intermediate (so that errors do not compound): the pasted
    frame is bit-exact given the device's u8 crop and soft mask, and the u8 crop is within 1 count of the oracle generator's."""
    from canonswap_b200 import synth, spec
    from canonswap_b200.modules import can_swapper
    from canonswap_b200.pipeline import FullLoopPipeline
    from oracle import canonswap_oracle as O
    w = dict(synth_w)
    w[spec.MOTION_NET] = synth.synth_motion_state_dict()
    sw = can_swapper(weights=w, device_id=0, max_batch=2)
    inp = synth.synth_inputs(2, 128, u8=True)
    sw.set_source_identity(inp["source_id"])
    H, W = 300, 420
    rng = np.random.default_rng(0)
    full = torch.from_numpy(rng.integers(0, 256, (2, H, W, 3), dtype=np.uint8)).pin_memory()
    pm = torch.cat([_face_mask(256, 256, seed=s)[0] for s in (1, 2)]).pin_memory()                 # [2,256,256]
    M = np.stack([np.array([[0.9, 0.1, 60.0 + 10 * i], [-0.1, 0.9, 30.0], [0, 0, 1]], dtype=np.float32) for i in range(2)])
    out = torch.empty_like(full).pin_memory()
    pipe = FullLoopPipeline(sw, net_hw=(128, 128), batch=2)
    assert pipe.run(inp["frames"].pin_memory(), pm, M, full, out) == 2
    # stage-wise re-computation on the device, then the oracle on the same intermediates
    eng = sw.engine((128, 128), 2)
    I_p, _ = sw.swap_frames(inp["frames"].cuda())
    soft, _ = pipe.soft_mask(pm.cuda()[:, None])
    for i in range(2):
        want = P.paste_back_frame(I_p[i].cpu().numpy(), soft[i, 0].cpu().numpy(), M[i], full[i].numpy())
        assert np.array_equal(out[i].numpy(), want), i
    kp = eng.keypoints(eng.motion(inp["frames"].cuda().permute(0, 3, 1, 2).float() / 255.0))
    ref = O.frame(synth_w, inp["frames"].permute(0, 3, 1, 2).float() / 255.0, kp["x_s"].cpu(), kp["x_can"].cpu(), inp["source_id"])["out"]
    ref_u8 = O.parse_output(ref)
    assert (ref_u8.int() - I_p.cpu().int()).abs().max().item() <= 1


@pytest.mark.gpu
def test_parse_mask_matches_torch_bit_for_bit():
    """LOOP A post-processing (reference can_swap_pipeline_e2e.py:183-190) fused on the device: labels and mask equal torch's
    interpolate -> argmax -> isin exactly, on random logits, on integer logits (exact arithmetic: real ties, first index wins),
    for the Segformer geometry (19 x 128 x 128 -> 512 x 512) and a ragged one; the mask feeds SoftErosion directly."""
    from canonswap_b200.engine import Engine
    eng = Engine(None, net_hw=(128, 128), max_batch=1, device=0)
    try:
        valid = torch.tensor(Engine.VALID_PARSE_LABELS, device="cuda")
        g = torch.Generator(device="cuda").manual_seed(3)
        cases = [(2, 19, 128, 128, 512, 512, "randn"), (1, 19, 128, 128, 512, 512, "int"), (2, 7, 33, 21, 100, 77, "randn"),
                 (1, 19, 64, 64, 512, 512, "int")]
        for (B, C, h, w, H, W, kind) in cases:
            if kind == "randn":
                lg = torch.randn(B, C, h, w, device="cuda", generator=g) * 4
            else:
                lg = torch.randint(-3, 4, (B, C, h, w), device="cuda", generator=g).float()
            up = torch.nn.functional.interpolate(lg, size=(H, W), mode="bilinear", align_corners=False)
            lab = up.argmax(dim=1)
            ref = torch.isin(lab, valid).to(torch.int)
            mask, labels = eng.parse_mask(lg, (H, W), want_labels=True)
            assert torch.equal(labels.long(), lab), (kind, (labels.long() != lab).sum().item())
            assert torch.equal(mask, ref.float())
        from canonswap_b200.pasteback import SoftErosion
        soft, hard = SoftErosion(21, 0.9, 3).bind(eng)(mask[:, None])
        assert soft.shape == (1, 1, 512, 512) and torch.isfinite(soft).all()
    finally:
        eng.close()
