"""Development probe (GPU box): relative error of the tcgen05 conv against fp64 torch when the ACTIVATIONS are scaled by
1e-4 .. 1e4 (weights O(1/sqrt(fan_in))): the split-fp16 operand format has fp16's exponent range."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from canonswap_b200.engine import Engine
eng = Engine(None, net_hw=(128, 128), max_batch=1, device=0)
g = torch.Generator(device="cuda").manual_seed(7)
for (Cin, Cout, k) in ((128, 256, 3), (512, 512, 3), (256, 512, 1)):
    x0 = torch.randn(2, 1, 32, 32, Cin, device="cuda", generator=g)
    w = torch.randn(Cout, Cin, 1, k, k, device="cuda", generator=g) / (Cin * k * k) ** 0.5
    pad = (0, k // 2, k // 2)
    for s in (1e-5, 1e-4, 1e-3, 1e-2, 1e-1, 1.0, 1e1, 1e2, 1e3, 1e4, 1e5):
        x = x0 * s
        ref = F.conv3d(x.permute(0, 4, 1, 2, 3).double(), w.double(), None, padding=pad).permute(0, 2, 3, 4, 1)
        outs = []
        for impl in (2, 5) if k == 3 else (2,):
            y = eng.test_conv(x, w, None, pad, impl=impl)
            e = (y.double() - ref)
            outs.append(f"impl{impl}: rms={(e.pow(2).mean().sqrt() / ref.abs().mean()).item():.2e} max={(e.abs().max() / ref.abs().max()).item():.1e}")
        print(f"{Cin}->{Cout} k{k} act scale {s:g}: " + "  ".join(outs), flush=True)
eng.close()
