"""The reference's own LOOP C text (src/can_swap_pipeline_e2e.py:223-283) and make_motion_template (:101-135), executed
unmodified through tests/ref_pipeline_harness.py:

  CPU  (not gpu): against an oracle-backed can_swapper -- proves the harness runs the reference text and that the oracle
       composition equals what the loop computes (I_p, the two debug decodes, the pasted frame);
  GPU: against the drop-in canonswap_b200.modules.can_swapper -- "inference_canswap.py calls it unchanged" for the part of
       the pipeline that touches the hot path.  The B200 loop's outputs are compared with the oracle loop fed the SAME
       keypoints (the image moves ~1e-2 per 1e-5 of keypoint shift on this fixture).
"""
import numpy as np
import pytest
import torch

import ref_pipeline_harness as H
from canonswap_b200 import synth
from oracle import canonswap_oracle as O

pytestmark = pytest.mark.skipif(H.reference_root() is None, reason="needs /root/reference or the oracle/_ref bundle")
NET = 128


def _np(v):
    return v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)     # dct2device turns the template into tensors in place


class _OracleWarp:
    def __init__(self, sd):
        self.sd = sd

    def warp(self, feature_3d, kp_source, kp_driving):
        return O.warp(self.sd, feature_3d, kp_source, kp_driving)

    def __call__(self, feature_3d, kp_driving=None, kp_source=None):
        return O.warping_forward(self.sd, feature_3d, kp_driving=kp_driving, kp_source=kp_source)


class OracleSwapper:
    """The can_swapper surface LOOP C uses, computed by the CPU oracle. kp_replay: per-frame (x_s, kp, scale) to return from
    get_kp_info / transform_keypoint instead of the oracle's own motion extractor."""
    device = "cpu"

    def __init__(self, W, kp_replay=None):
        self.W, self.kp_replay, self._i = W, kp_replay, 0
        self.warping_module = _OracleWarp(W["warping_module"])
        self.swap_module = lambda f, sid: O.swap_module(W["transfer"], f, sid.expand(f.shape[0], -1))
        self.refine_module = lambda f: O.refine_module(W["refine"], f)

    def prepare_videos(self, imgs):
        return O.prepare_videos(torch.from_numpy(np.array(imgs)))

    def get_kp_info(self, x):
        if self.kp_replay is not None:
            r = self.kp_replay[self._i]
            self._i += 1
            return {k: torch.from_numpy(np.array(v)) for k, v in r.items()}
        info = O.motion_extractor(self.W["motion_extractor"], x)
        bs = x.shape[0]
        deg = {k: O.headpose_pred_to_degree(info[k])[:, None] for k in ("pitch", "yaw", "roll")}
        return {**info, **deg, "kp": info["kp"].reshape(bs, -1, 3), "exp": info["exp"].reshape(bs, -1, 3)}

    def transform_keypoint(self, info):
        if "x_s" in info:
            return info["x_s"]
        return O.transform_keypoint({**info, "kp": info["kp"].reshape(info["kp"].shape[0], -1)})

    def extract_feature_3d(self, x):
        return O.appearance_feature_extractor(self.W["appearance_feature_extractor"], x)

    def conv_decode(self, out, occ=None):
        return O.spade_decoder(self.W["spade_generator"], O.warp_out(self.W["warping_module"], out, occ))

    def warp_decode(self, f, kp_source, kp_driving):
        r = self.warping_module(f, kp_source=kp_source, kp_driving=kp_driving)
        r["out"] = O.spade_decoder(self.W["spade_generator"], r["out"])
        return r

    def parse_output(self, out):
        return O.parse_output(out).numpy()

    def prepare_source(self, img):
        return torch.from_numpy(np.clip(img[None].astype(np.float32) / 255.0, 0, 1)).permute(0, 3, 1, 2)


class _CpuSoftErosion:
    """SoftErosion(21, 0.9, 3) of the reference (src/utils/crop.py:21-47), from the file itself."""

    def __init__(self):
        mod, _ = H.load_pipeline_module(OracleSwapper)
        self.m = mod.SoftErosion(kernel_size=21, threshold=0.9, iterations=3)

    def __call__(self, x):
        with torch.no_grad():
            return self.m(x)


def _inputs(T):
    W = synth.synth_weights(with_motion=True)
    clip = synth.synth_inputs(T, NET, u8=True)
    g = torch.Generator().manual_seed(5)
    masks = (torch.rand(T, 2 * NET, 2 * NET, generator=g) > 0.35).float()
    full = [np.random.RandomState(i).randint(0, 256, (300, 400, 3)).astype(np.uint8) for i in range(T)]
    M = [np.array([[1.1, -0.1, 60.0], [0.1, 1.1, 20.0], [0, 0, 1]], np.float32)] * T
    return W, clip, masks, full, M


def _oracle_expected(W, clip, kp):
    """What LOOP C computes, from the oracle composition fed the template's keypoints."""
    fr = clip["frames"].permute(0, 3, 1, 2).float() / 255.0
    x_t = torch.cat([torch.from_numpy(_np(m["x_s"])) for m in kp])
    x_can = torch.cat([torch.from_numpy(_np(m["scale"]))[..., None] * torch.from_numpy(_np(m["kp"])) for m in kp])
    r = O.frame(W, fr, x_t, x_can, clip["source_id"], debug_decodes=True)
    return {"I_p": O.parse_output(r["out"]).numpy(), "rec_can": O.parse_output(r["rec_can"]).numpy(),
            "I_can": O.parse_output(r["swap_can"]).numpy()}


def test_reference_loop_text_runs_on_the_oracle_swapper():
    T = 1
    W, clip, masks, full, M = _inputs(T)
    sw = OracleSwapper(W)
    out = H.run_reference_loop(sw, _CpuSoftErosion(), [f.numpy() for f in clip["frames"]], clip["source_id"], masks, M, full)
    assert len(out["I_p_lst"]) == T and len(out["I_p_pstbk_lst"]) == T
    exp = _oracle_expected(W, clip, out["template"]["motion"])
    for i in range(T):
        assert out["I_p_lst"][i].shape == (2 * NET, 2 * NET, 3) and out["I_p_lst"][i].dtype == np.uint8
        for got, key in ((out["I_p_lst"][i], "I_p"), (out["rec_can_lst"][i], "rec_can"), (out["I_can_lst"][i], "I_can")):
            assert np.abs(got.astype(int) - exp[key][i].astype(int)).max() <= 1, key      # B=1 loop vs B=T batch: fp32 order only
        assert out["I_p_pstbk_lst"][i].shape == full[i].shape
    # the LOOP C text that was executed is the reference's, not a restatement
    src = H.loop_c_source(H.load_pipeline_module(OracleSwapper)[1])
    for needle in ("extract_feature_3d(I_s)", "warping_module.warp(f_s, x_t, x_can)", "swap_module(f_can, source_id)",
                   "refine_module(f_can_swap)", "warp_decode(f_can_swap, x_can, x_t)", "paste_back(I_p_i"):
        assert needle in src


@pytest.mark.gpu
def test_reference_loop_text_runs_on_the_b200_swapper():
    from canonswap_b200.modules import can_swapper
    from canonswap_b200.pasteback import SoftErosion
    T = 2
    W, clip, masks, full, M = _inputs(T)
    sw = can_swapper(weights=W, device_id=0, max_batch=1)
    sw.input_shape = (NET, NET)
    soft = SoftErosion(kernel_size=21, threshold=0.9, iterations=3).bind(sw.engine((NET, NET), 1))
    sid = clip["source_id"].cuda()
    out = H.run_reference_loop(sw, soft, [f.numpy() for f in clip["frames"]], sid, masks.cuda(), M, full)
    kp = out["template"]["motion"]
    exp = _oracle_expected(W, clip, kp)
    # keypoints from the B200 motion extractor against the oracle's
    mk = O.motion_keypoints(W["motion_extractor"], clip["frames"].permute(0, 3, 1, 2).float() / 255.0)
    x_s = np.concatenate([_np(m["x_s"]) for m in kp])
    assert np.abs(x_s - mk["x_t"].numpy()).max() <= 1e-4
    for i in range(T):
        for got, key in ((out["I_p_lst"][i], "I_p"), (out["rec_can_lst"][i], "rec_can"), (out["I_can_lst"][i], "I_can")):
            d = np.abs(got.astype(int) - exp[key][i].astype(int))
            assert d.max() <= 1 and (d > 0).mean() < 0.02, (key, d.max(), (d > 0).mean())
    # the pasted frames: the same loop text over the oracle swapper replaying the SAME keypoints (CPU SoftErosion, cv2 paste-back)
    replay = [{"x_s": _np(m["x_s"]), "kp": _np(m["kp"]), "scale": _np(m["scale"]), "exp": _np(m["exp"]), "t": _np(m["t"]),
               "pitch": np.zeros((1, 1), np.float32), "yaw": np.zeros((1, 1), np.float32), "roll": np.zeros((1, 1), np.float32)} for m in kp]
    ref = H.run_reference_loop(OracleSwapper(W, kp_replay=replay), _CpuSoftErosion(), [f.numpy() for f in clip["frames"]],
                               clip["source_id"], masks, M, full)
    for i in range(T):
        d = np.abs(out["I_p_pstbk_lst"][i].astype(int) - ref["I_p_pstbk_lst"][i].astype(int))
        assert d.max() <= 1 and (d > 0).mean() < 0.02, (d.max(), (d > 0).mean())


# ---- the video-to-image pipeline -------------------------------------------------------------
def _v2i_inputs(T):
    W = synth.synth_weights(with_motion=True)
    clip = synth.synth_inputs(T + 1, NET, u8=True)                  # frame 0 = the source image, 1.. = the driving video
    full = np.random.RandomState(3).randint(0, 256, (300, 400, 3)).astype(np.uint8)
    M = np.array([[1.1, -0.1, 60.0], [0.1, 1.1, 20.0], [0, 0, 1]], np.float32)
    mask = np.random.RandomState(4).rand(300, 400, 3).astype(np.float32)
    return W, clip, full, M, mask


def test_v2i_reference_text_matches_the_oracle_composition():
    """Pins oracle.v2i_source_state / v2i_frames: the reference's own v2i texts (execute_face_canonical's hot statements, the
    frame loop incl. its `i == 0` block) over the oracle-backed swapper give the images the oracle composition gives."""
    T = 1
    W, clip, full, M, mask = _v2i_inputs(T)
    sw = OracleSwapper(W)
    out = H.run_reference_v2i(sw, clip["frames"][0].numpy(), clip["source_id"], [f.numpy() for f in clip["frames"][1:]], full, M, mask)
    I_s = clip["frames"][:1].permute(0, 3, 1, 2).float() / 255.0
    st = O.v2i_source_state(W, I_s, clip["source_id"])
    delta = torch.cat([torch.from_numpy(_np(m["exp"])) for m in out["template"]["motion"]])
    img, _ = O.v2i_frames(W, st, delta)
    exp = O.parse_output(img).numpy()
    assert (st["f_s_can"] - out["f_s_can"]).abs().max().item() <= 1e-5
    assert np.abs(out["I_can_lst"][0].astype(int) - O.parse_output(st["swap_can"]).numpy()[0].astype(int)).max() <= 1
    for i in range(T):
        assert np.abs(out["I_p_lst"][i].astype(int) - exp[i].astype(int)).max() <= 1
    src = H.v2i_loop_source(H.load_pipeline_module(OracleSwapper, "can_swap_pipeline_v2i")[1])
    for needle in ("swap_module(f_s_can, driving_id)", "x_swap_info['kp'] @ R_swap + delta_t", "warp_decode(f_swap_can_2, x_swap, x_t_2)"):
        assert needle in src


@pytest.mark.gpu
def test_v2i_pipeline_b200_vs_oracle():
    """V2IPipeline (prepare once, resident appearance volume, one cs_frame per batch) against the oracle composition fed the
    same driving expressions; the state against the oracle's (tight), the frames at the image bar."""
    from canonswap_b200.modules import can_swapper
    from canonswap_b200.pipeline import V2IPipeline
    T = 3
    W, clip, full, M, mask = _v2i_inputs(T)
    sw = can_swapper(weights=W, device_id=0, max_batch=4)
    sw.input_shape = (NET, NET)
    A = O.V2I_ANIMATE_HW[0]                                          # the reference animates at 256 whatever the source size
    pipe = V2IPipeline(sw, net_hw=(A, A), batch=4)
    I_s = (clip["frames"][:1].permute(0, 3, 1, 2).float() / 255.0).cuda()
    st = pipe.prepare(I_s, clip["source_id"].cuda())
    ost = O.v2i_source_state(W, I_s.cpu(), clip["source_id"])
    assert (st["x_swap"].cpu() - ost["x_swap"]).abs().max().item() <= 5e-4       # keypoints of a decoded image: image noise x M
    assert (st["R_swap"].cpu() - ost["R_swap"]).abs().max().item() <= 1e-5
    assert (st["swap_can"].cpu() - ost["swap_can"]).abs().max().item() <= 1e-3
    g = torch.Generator().manual_seed(9)
    delta = (0.02 * torch.randn(T, 21, 3, generator=g)).pin_memory()
    out = torch.empty(T, 2 * A, 2 * A, 3, dtype=torch.uint8).pin_memory()
    assert pipe.run(delta, out) == T
    # the oracle frame body fed the B200 state (the image moves ~1e-2 per 1e-5 of keypoint shift on this fixture)
    ost_dev = dict(ost, **{k: st[k].cpu() for k in ("feature", "x_swap", "kp_swap", "R_swap", "t_swap", "scale_swap")})
    ost_dev["feature"] = O.appearance_feature_extractor(W["appearance_feature_extractor"],
                                                        torch.nn.functional.interpolate(st["swap_can"].cpu(), size=(A, A), mode="bilinear",
                                                                                        align_corners=False))
    img, _ = O.v2i_frames(W, ost_dev, delta)
    exp = O.parse_output(img).numpy()
    d = np.abs(out.numpy().astype(int) - exp.astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 0.02, (d.max(), (d > 0).mean())
    # sharded: ranks 0 / 1 of a world of 2 produce exactly the single-rank frames (no process group needed: state is local)
    for r in range(2):
        o2 = torch.zeros_like(out)
        pipe.run(delta, o2, rank=r, world=2)
        for i in range(r, T, 2):
            assert torch.equal(o2[i], out[i])
