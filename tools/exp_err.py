"""Development probe (GPU box): end-to-end max|d| vs the CPU oracle and step time for option settings.
    python tools/exp_err.py NET "11=0" "11=128,8=0" ...   (each argument = one setting: comma-separated option=value)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from canonswap_b200 import synth, _lib
from canonswap_b200.engine import Engine
from oracle import canonswap_oracle as O

net = int(sys.argv[1])
settings = sys.argv[2:] or [""]
B = 2 if net <= 128 else 1
W = synth.synth_weights()
inp = synth.synth_inputs(B, net, seed=11)
ref = O.frame(W, inp["frames"], inp["x_t"], inp["x_can"], inp["source_id"], debug_decodes=False)["out"]
cu = {k: v.cuda() for k, v in inp.items()}
for st in settings:
    kvs = [x.split("=") for x in st.split(",") if x]
    eng = Engine(W, net_hw=(net, net), max_batch=B, device=0, options={int(a): int(b) for a, b in kvs})
    eng.set_identity(cu["source_id"])
    o = torch.empty(B, 3, 2 * net, 2 * net, device="cuda")
    eng.frame(cu["frames"], cu["x_t"], cu["x_can"], out_f32=o)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        eng.frame(cu["frames"], cu["x_t"], cu["x_can"], out_f32=o)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 3 * 1e3
    d = (o.cpu() - ref).abs().max().item()
    print(f"net={net} [{st}] max|d|={d:.3e} step={ms:.2f} ms (B={B})", flush=True)
    eng.close()
