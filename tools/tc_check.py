"""Development probe (GPU box): tcgen05 conv vs torch on many shapes, errors + per-launch timing.
Prints one line per case; never asserts, so one run shows every shape."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from canonswap_b200 import synth, _lib
from canonswap_b200.engine import Engine

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

CASES = [
    # B, D, H, W, Cin, Cout, k, pad, time?
    (1, 1, 8, 16, 64, 64, (1, 1, 1), (0, 0, 0), 0),
    (1, 1, 8, 16, 64, 64, (1, 3, 3), (0, 1, 1), 0),
    (2, 1, 8, 8, 64, 128, (1, 3, 3), (0, 1, 1), 0),
    (1, 1, 8, 8, 256, 512, (1, 1, 1), (0, 0, 0), 0),
    (1, 16, 8, 8, 32, 32, (3, 3, 3), (1, 1, 1), 0),
    (1, 16, 4, 4, 110, 64, (3, 3, 3), (1, 1, 1), 0),
    (1, 16, 8, 8, 142, 22, (7, 7, 7), (3, 3, 3), 0),
    (1, 16, 2, 2, 64, 48, (3, 3, 3), (1, 1, 1), 0),
    (3, 16, 1, 1, 512, 1024, (3, 3, 3), (1, 1, 1), 0),
    (1, 1, 8, 8, 64, 12, (1, 3, 3), (0, 1, 1), 0),
    (1, 1, 5, 7, 20, 33, (1, 3, 3), (0, 1, 1), 0),
    (1, 1, 24, 40, 48, 272, (1, 3, 3), (0, 1, 1), 0),
    # real shapes, timed
    (8, 1, 64, 64, 512, 512, (1, 3, 3), (0, 1, 1), 1),
    (8, 1, 64, 64, 512, 1024, (1, 3, 3), (0, 1, 1), 1),
    (8, 1, 64, 64, 256, 128, (1, 3, 3), (0, 1, 1), 1),
    (8, 1, 64, 64, 128, 1024, (1, 3, 3), (0, 1, 1), 1),
    (8, 16, 64, 64, 32, 32, (3, 3, 3), (1, 1, 1), 1),
    (2, 16, 64, 64, 142, 142, (3, 3, 3), (1, 1, 1), 1),
    (2, 16, 64, 64, 142, 22, (7, 7, 7), (3, 3, 3), 1),
    (8, 1, 256, 256, 64, 64, (1, 3, 3), (0, 1, 1), 1),
    (8, 1, 256, 256, 128, 64, (1, 3, 3), (0, 1, 1), 1),
    (8, 1, 128, 128, 256, 256, (1, 3, 3), (0, 1, 1), 1),
    (8, 1, 256, 256, 128, 512, (1, 3, 3), (0, 1, 1), 1),
    (8, 1, 64, 64, 256, 512, (1, 3, 3), (0, 1, 1), 1),
    # the depth-stacked 32->32 3x3x3 kernel (impl 4)
    (1, 16, 16, 8, 32, 32, (3, 3, 3), (1, 1, 1), 5),
    (2, 16, 24, 40, 32, 32, (3, 3, 3), (1, 1, 1), 5),
    (8, 16, 64, 64, 32, 32, (3, 3, 3), (1, 1, 1), 6),
    # the depth-stacked 7x7x7 kernel (impl 3)
    (1, 16, 16, 8, 142, 22, (7, 7, 7), (3, 3, 3), 3),
    (2, 16, 32, 32, 142, 22, (7, 7, 7), (3, 3, 3), 3),
    (1, 16, 24, 40, 70, 24, (7, 7, 7), (3, 3, 3), 3),
    (2, 16, 64, 64, 142, 22, (7, 7, 7), (3, 3, 3), 4),
    (8, 16, 64, 64, 142, 22, (7, 7, 7), (3, 3, 3), 4),
]


def main():
    passes = [3]
    comps = [int(a) for a in sys.argv[1:] if a[0].isdigit()] or [120]
    pair = 0 if "nopair" in sys.argv[1:] else 1
    dbuf = 0 if "nodbuf" in sys.argv[1:] else 1
    bnmax = 128 if "bn128" in sys.argv[1:] else 0
    eng = Engine(None, net_hw=(128, 128), max_batch=1, device=0)
    g = torch.Generator(device="cuda").manual_seed(7)
    for comp in comps:
      for np_ in passes:
        eng.set_option(_lib.CS_OPT_TC_COMP, comp)
        eng.set_option(_lib.CS_OPT_TC_PAIR, pair)
        eng.set_option(_lib.CS_OPT_TC_DOUBLE_BUFFER, dbuf)
        eng.set_option(_lib.CS_OPT_TC_BN_MAX, bnmax)
        for (B, D, H, Wd, Cin, Cout, k, pad, timed) in CASES:
            x = torch.randn(B, D, H, Wd, Cin, device="cuda", generator=g)
            w = torch.randn(Cout, Cin, *k, device="cuda", generator=g) / (Cin * k[0] * k[1] * k[2]) ** 0.5
            b = torch.randn(Cout, device="cuda", generator=g)
            tag = f"comp={comp} pair={pair} dbuf={dbuf} bn={bnmax} B{B} D{D} {H}x{Wd} {Cin}->{Cout} k{k}"
            impl, act = (4, 2) if timed >= 5 else ((3, 0) if timed >= 3 else (2, 2))
            timed = timed in (1, 4, 6)
            try:
                eng.profile(True)
                y = eng.test_conv(x, w, b, pad, act=act, slope=0.2, impl=impl)
                if timed:
                    for _ in range(3):
                        y = eng.test_conv(x, w, b, pad, act=act, slope=0.2, impl=impl)
                torch.cuda.synchronize()
                fam = eng.profile_read()["conv_tcgen05"]
                eng.profile(False)
                xr = x.permute(0, 4, 1, 2, 3).contiguous()
                if timed:
                    ref = F.conv3d(xr, w, b, padding=pad).permute(0, 2, 3, 4, 1)
                else:
                    ref = F.conv3d(xr.double(), w.double(), b.double(), padding=pad).permute(0, 2, 3, 4, 1).float()
                if act == 2:
                    ref = F.leaky_relu(ref, 0.2)
                big = ref.abs() > 1.0
                bias = (((y - ref) * ref.sign())[big].double().mean() / ref[big].abs().double().mean()).item() if big.any() else 0.0
                rms = ((y - ref)[big].double().pow(2).mean().sqrt() / ref[big].abs().double().mean()).item() if big.any() else 0.0
                tag += f" bias={bias:+.2e} rms={rms:.2e}"
                err = (y - ref).abs().max().item()
                scale = ref.abs().max().item()
                ms = fam["ms"] / max(1, fam["launches"])
                tf = fam["flops"] / max(1, fam["launches"]) / (ms * 1e9) if ms > 0 else 0
                print(f"{tag}: max|d|={err:.3e} (ref max {scale:.2f}) {ms:.3f} ms {tf:.1f} TF/s-equiv", flush=True)
            except Exception as e:
                print(f"{tag}: FAILED {type(e).__name__}: {e}", flush=True)
                if "CUDA" in str(e) or "launch" in str(e) or "illegal" in str(e) or "unspecified" in str(e):
                    raise
    eng.close()


if __name__ == "__main__":
    main()
