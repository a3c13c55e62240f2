"""GPU tests, kernel level: the C-ABI test entry points against the matching torch.nn.functional op
(fp32, TF32 off) on random tensors including edge shapes."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng(synth_w):
    from canonswap_b200.engine import Engine
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    e = Engine(synth_w, net_hw=(128, 128), max_batch=2, device=0)
    yield e
    e.close()


def _ref_conv(x_cl, w, b, pad):
    x = x_cl.permute(0, 4, 1, 2, 3).contiguous().double()
    y = F.conv3d(x, w.double(), None if b is None else b.double(), padding=pad)
    return y.permute(0, 2, 3, 4, 1).float()


CONV_CASES = [
    # B, D, H, W, Cin, Cout, k(d,h,w), pad
    (1, 1, 16, 16, 3, 64, (1, 3, 3), (0, 1, 1)),      # F.first
    (2, 1, 8, 8, 64, 128, (1, 3, 3), (0, 1, 1)),
    (1, 1, 8, 8, 256, 512, (1, 1, 1), (0, 0, 0)),     # 1x1
    (1, 16, 8, 8, 32, 32, (3, 3, 3), (1, 1, 1)),      # ResBlock3d
    (1, 16, 4, 4, 110, 64, (3, 3, 3), (1, 1, 1)),     # hourglass conv0, odd Cin
    (1, 16, 8, 8, 142, 22, (7, 7, 7), (3, 3, 3)),     # mask conv, odd Cin / Cout
    (1, 16, 2, 2, 64, 48, (3, 3, 3), (1, 1, 1)),      # tiny spatial
    (1, 1, 8, 8, 512, 1, (1, 3, 3), (0, 1, 1)),       # mask_conv (Cout 1)
    (1, 16, 8, 8, 142, 1, (16, 7, 7), (0, 3, 3)),     # occlusion as conv3d with full-depth kernel
    (1, 1, 8, 8, 64, 12, (1, 3, 3), (0, 1, 1)),       # conv_img
    (1, 1, 5, 7, 20, 33, (1, 3, 3), (0, 1, 1)),       # ragged everything
]


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("act", [0, 2])
def test_conv_simt_matches_torch(eng, case, act):
    B, D, H, W, Cin, Cout, k, pad = case
    if Cout == 1 and act == 2:
        pytest.skip("Cout=1 kernel supports none/sigmoid only")
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(B, D, H, W, Cin, device="cuda", generator=g)
    w = torch.randn(Cout, Cin, *k, device="cuda", generator=g) / (Cin * k[0] * k[1] * k[2]) ** 0.5
    b = torch.randn(Cout, device="cuda", generator=g)
    y = eng.test_conv(x, w, b, pad, act=act, slope=0.2, impl=1)
    ref = _ref_conv(x, w, b, pad)
    if act == 2:
        ref = F.leaky_relu(ref, 0.2)
    assert y.shape == ref.shape
    assert (y - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item())


TC_CASES = [c for c in CONV_CASES if c[4] >= 16] + [
    (3, 16, 1, 1, 512, 1024, (3, 3, 3), (1, 1, 1)),   # deepest hourglass level at net 128: 1x1 spatial, batch-tiled box
    (1, 1, 24, 40, 48, 272, (1, 3, 3), (0, 1, 1)),    # non-power-of-two extents, two N tiles of 144
    (2, 1, 64, 64, 512, 512, (1, 3, 3), (0, 1, 1)),   # the most-used shape
    (2, 1, 64, 48, 3, 64, (1, 3, 3), (0, 1, 1)),      # RGB input of the first conv: 3 channels padded to one K step
]


@pytest.mark.parametrize("case", TC_CASES)
@pytest.mark.parametrize("act", [0, 2])
def test_conv_tcgen05_matches_torch(eng, case, act):
    """The tcgen05 split-bf16 implicit GEMM (3 MMA passes, fp32 accumulate in TMEM) against fp64 torch:
    fp32-grade agreement, far inside the 1e-3 parity bar."""
    B, D, H, W, Cin, Cout, k, pad = case
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(B, D, H, W, Cin, device="cuda", generator=g)
    w = torch.randn(Cout, Cin, *k, device="cuda", generator=g) / (Cin * k[0] * k[1] * k[2]) ** 0.5
    b = torch.randn(Cout, device="cuda", generator=g)
    y = eng.test_conv(x, w, b, pad, act=act, slope=0.2, impl=2)
    ref = _ref_conv(x, w, b, pad)
    if act == 2:
        ref = F.leaky_relu(ref, 0.2)
    assert y.shape == ref.shape
    assert (y - ref).abs().max().item() <= 5e-5 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("case", [(1, 16, 16, 8, 142, 22), (2, 16, 32, 32, 142, 22), (1, 16, 24, 40, 70, 24)])
def test_conv7_depth_stacked_matches_torch(eng, case):
    """The dedicated 7x7x7 mask-conv kernel (depth-stacked N, kh-split partial sums) against fp64 torch."""
    B, D, H, W, Cin, Cout = case
    g = torch.Generator(device="cuda").manual_seed(13)
    x = torch.randn(B, D, H, W, Cin, device="cuda", generator=g)
    w = torch.randn(Cout, Cin, 7, 7, 7, device="cuda", generator=g) / (Cin * 343) ** 0.5
    b = torch.randn(Cout, device="cuda", generator=g)
    y = eng.test_conv(x, w, b, (3, 3, 3), impl=3)
    ref = _ref_conv(x, w, b, (3, 3, 3))
    assert y.shape == ref.shape
    assert (y - ref).abs().max().item() <= 5e-5 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("case", [(1, 32, 32, 128, 256), (2, 64, 64, 512, 512), (1, 16, 32, 256, 256), (1, 24, 40, 128, 256)])
@pytest.mark.parametrize("act", [0, 2])
def test_conv_winograd_matches_torch(eng, case, act):
    """Winograd F(2x2,3x3) form of a 3x3 2-D conv (input transform -> 16 GEMMs with depth-dependent weights on the persistent
    tcgen05 kernel -> output transform) against fp64 torch, same bar as the direct tcgen05 conv."""
    B, H, W, Cin, Cout = case
    g = torch.Generator(device="cuda").manual_seed(21)
    x = torch.randn(B, 1, H, W, Cin, device="cuda", generator=g)
    w = torch.randn(Cout, Cin, 1, 3, 3, device="cuda", generator=g) / (Cin * 9) ** 0.5
    b = torch.randn(Cout, device="cuda", generator=g)
    y = eng.test_conv(x, w, b, (0, 1, 1), act=act, slope=0.2, impl=5)
    ref = _ref_conv(x, w, b, (0, 1, 1))
    if act == 2:
        ref = F.leaky_relu(ref, 0.2)
    assert y.shape == ref.shape
    assert (y - ref).abs().max().item() <= 5e-5 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("case", [(1, 16, 8), (2, 24, 40), (2, 32, 32)])
@pytest.mark.parametrize("act", [0, 2])
def test_conv3_depth_stacked_matches_torch(eng, case, act):
    """The dedicated 32->32 3x3x3 volume-conv kernel (depth-stacked N, weights resident in smem) against fp64 torch."""
    B, H, W = case
    g = torch.Generator(device="cuda").manual_seed(14)
    x = torch.randn(B, 16, H, W, 32, device="cuda", generator=g)
    w = torch.randn(32, 32, 3, 3, 3, device="cuda", generator=g) / (32 * 27) ** 0.5
    b = torch.randn(32, device="cuda", generator=g)
    y = eng.test_conv(x, w, b, (1, 1, 1), act=act, slope=0.2, impl=4)
    ref = _ref_conv(x, w, b, (1, 1, 1))
    if act == 2:
        ref = F.leaky_relu(ref, 0.2)
    assert y.shape == ref.shape
    assert (y - ref).abs().max().item() <= 5e-5 * max(1.0, ref.abs().max().item())


def test_conv_tcgen05_epilogue_residual_mult_strided(eng):
    """Engine-level check of the fused epilogue is in test_gpu_stages (warp_out: x occlusion, resblocks:
    + residual); here: the same conv through both implementations must agree to fp32 round-off."""
    g = torch.Generator(device="cuda").manual_seed(12)
    x = torch.randn(2, 1, 32, 32, 256, device="cuda", generator=g)
    w = torch.randn(256, 256, 1, 3, 3, device="cuda", generator=g) / 48.0
    b = torch.randn(256, device="cuda", generator=g)
    y_tc = eng.test_conv(x, w, b, (0, 1, 1), act=1, impl=2)
    y_simt = eng.test_conv(x, w, b, (0, 1, 1), act=1, impl=1)
    assert (y_tc - y_simt).abs().max().item() <= 5e-5 * max(1.0, y_simt.abs().max().item())


def test_conv_sigmoid_cout1(eng):
    g = torch.Generator(device="cuda").manual_seed(8)
    x = torch.randn(2, 1, 8, 8, 512, device="cuda", generator=g)
    w = torch.randn(1, 512, 1, 3, 3, device="cuda", generator=g) / 68.0
    b = torch.randn(1, device="cuda", generator=g)
    y = eng.test_conv(x, w, b, (0, 1, 1), act=3, impl=1)
    ref = torch.sigmoid(_ref_conv(x, w, b, (0, 1, 1)))
    assert (y - ref).abs().max().item() <= 1e-5


@pytest.mark.parametrize("hw", [(8, 8), (32, 32), (16, 24)])
def test_grid_sample3d_matches_torch(eng, hw):
    """5-D trilinear, zeros padding, align_corners=False (reference warping_network.py:46-47), including
    out-of-range and exactly-on-border coordinates."""
    H, W = hw
    g = torch.Generator(device="cuda").manual_seed(9)
    inp = torch.randn(2, 32, 16, H, W, device="cuda", generator=g)
    grid = torch.rand(2, 16, H, W, 3, device="cuda", generator=g) * 2.6 - 1.3
    grid[0, 0, 0, 0] = torch.tensor([-1.0, -1.0, -1.0], device="cuda")
    grid[0, 0, 0, 1] = torch.tensor([1.0, 1.0, 1.0], device="cuda")
    grid[0, 0, 0, 2] = torch.tensor([0.0, 0.0, 0.0], device="cuda")
    grid[0, 0, 0, 3] = torch.tensor([5.0, -5.0, 0.3], device="cuda")
    out = eng.test_grid_sample3d(inp, grid)
    ref = F.grid_sample(inp, grid, mode="bilinear", padding_mode="zeros", align_corners=False)
    assert (out - ref).abs().max().item() <= 2e-5


def test_grid_sample3d_identity_grid(eng):
    """The reference's identity grid (align_corners=True style, util.py:41-58) sampled with
    align_corners=False is NOT the identity map; both sides must agree on that."""
    from oracle import canonswap_oracle as O
    inp = torch.randn(1, 32, 16, 16, 16, device="cuda")
    grid = O.make_coordinate_grid(16, 16, 16, device="cuda")[None].contiguous()
    out = eng.test_grid_sample3d(inp, grid)
    ref = F.grid_sample(inp, grid, align_corners=False)
    assert (out - ref).abs().max().item() <= 2e-5


@pytest.mark.parametrize("shape", [(2, 32, 16 * 32 * 32), (2, 512, 32 * 32), (1, 64, 128 * 128), (2, 256, 1000)])
def test_instance_stats(eng, shape):
    B, C, S = shape
    x = torch.randn(B, C, S, device="cuda") * 3 + 1.5
    mean, rstd = eng.test_instance_stats(x, eps=1e-5)
    xd = x.double()
    rm = xd.mean(-1)
    rv = xd.var(-1, unbiased=False)
    assert (mean.double() - rm).abs().max().item() <= 1e-5
    assert ((rstd.double() - 1 / torch.sqrt(rv + 1e-5)).abs() * torch.sqrt(rv + 1e-5)).max().item() <= 1e-5


def test_bad_arguments_fail_loudly(eng):
    from canonswap_b200.engine import CanonSwapError
    with pytest.raises(ValueError):
        eng.appearance(torch.zeros(1, 3, 64, 64, device="cuda"))          # wrong resolution for this ctx
    with pytest.raises(ValueError):
        eng.appearance(torch.zeros(3, 3, 128, 128, device="cuda"))        # batch > max_batch
    with pytest.raises(ValueError):
        eng.appearance(torch.zeros(1, 3, 128, 128))                       # CPU tensor: no fallback
    with pytest.raises(CanonSwapError):
        eng.test_conv(torch.zeros(1, 1, 4, 4, 8, device="cuda"), torch.zeros(8, 8, 1, 3, 3, device="cuda"), None,
                      (0, 1, 1), impl=2 if not _tc_any(eng) else 7)


def _tc_any(eng):
    try:
        eng.test_conv(torch.zeros(1, 1, 8, 16, 64, device="cuda"), torch.zeros(64, 64, 1, 3, 3, device="cuda"), None,
                      (0, 1, 1), impl=2)
        return True
    except Exception:
        return False


def test_activation_prescale_restores_precision_of_small_activations():
    """The split-fp16 operand has fp16's exponent range: with activations of magnitude 1e-3 the lo halves are subnormal and the
    conv's relative error grows 50x; the per-conv power-of-two activation pre-scale (ConvW::amul, chosen by cs_calibrate) brings it
    back.  CS_OPT_TEST_AMUL applies a given scale to the test conv's operand."""
    from canonswap_b200 import _lib
    from canonswap_b200.engine import Engine
    e = Engine(None, net_hw=(128, 128), max_batch=1, device=0)
    try:
        g = torch.Generator(device="cuda").manual_seed(7)
        x = torch.randn(2, 1, 32, 32, 256, device="cuda", generator=g) * 1e-3
        w = torch.randn(256, 256, 1, 3, 3, device="cuda", generator=g) / (256 * 9) ** 0.5
        ref = _ref_conv(x, w, None, (0, 1, 1)).double()
        rms = {}
        for lg in (0, 10):
            e.set_option(_lib.CS_OPT_TEST_AMUL, lg)
            for impl in (2, 5):
                y = e.test_conv(x, w, None, (0, 1, 1), impl=impl)
                rms[(lg, impl)] = ((y.double() - ref).pow(2).mean().sqrt() / ref.abs().mean()).item()
        print("relative rms error, activations ~1e-3:", {k: f"{v:.1e}" for k, v in rms.items()})
        for impl in (2, 5):
            assert rms[(0, impl)] > 5e-6                      # the unscaled operand really loses precision ...
            assert rms[(10, impl)] < 1.5e-6                   # ... and the pre-scale restores it
    finally:
        e.close()


@pytest.mark.parametrize("case", [(1, 16, 8, 8, 128, 32, 3), (2, 16, 4, 4, 256, 64, 3), (1, 16, 2, 2, 512, 128, 3), (2, 1, 16, 16, 256, 128, 1),
                                  (1, 16, 6, 10, 96, 48, 3), (1, 16, 4, 4, 128, 512, 3)])
@pytest.mark.parametrize("act", [0, 1])
def test_conv_phase_form_matches_upsample_then_conv(eng, case, act):
    """The hourglass decoder convs read their input nearest-upsampled (1,2,2) (reference util.py:142-143).  Phase form: the conv
    runs on the LOW-resolution operand, one N tile per output phase (a, b) with the taps that fall on the same source pixel
    pre-summed -- 2 x 2 instead of 3 x 3 in-plane taps, no upsampled operand.  Against fp64 torch interpolate + conv3d."""
    B, D, H, W, Cin, Cout, KD = case
    g = torch.Generator(device="cuda").manual_seed(31)
    x = torch.randn(B, D, H, W, Cin, device="cuda", generator=g)
    w = torch.randn(Cout, Cin, KD, 3, 3, device="cuda", generator=g) / (Cin * KD * 9) ** 0.5
    b = torch.randn(Cout, device="cuda", generator=g)
    y = eng.test_conv_up2(x, w, b, act=act)
    xu = F.interpolate(x.permute(0, 4, 1, 2, 3).double(), scale_factor=(1, 2, 2), mode="nearest")
    ref = F.conv3d(xu, w.double(), b.double(), padding=(KD // 2, 1, 1)).permute(0, 2, 3, 4, 1).float()
    if act == 1:
        ref = F.relu(ref)
    assert y.shape == ref.shape
    assert (y - ref).abs().max().item() <= 5e-5 * max(1.0, ref.abs().max().item())
