"""CPU tests: the C-ABI library loads and exports every declared symbol, fails loudly without a GPU,
the host-side mirror keeps the reference state_dict layout, and the frame-sharding logic is exact
(incl. a world_size-2 gloo run)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from canonswap_b200 import _build, _lib
    _build.build()
    return _lib.load()


def test_library_exports_every_header_symbol(lib):
    from canonswap_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "canonswap_b200.h")).read()
    declared = re.findall(r"^CS_API [^;(]*?\b(cs_[a-z0-9_]+)\(", hdr, flags=re.M)
    assert sorted(declared) == sorted(_lib.SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_no_gpu_means_loud_failure_not_fallback(lib):
    import ctypes as C
    ctx = C.c_void_p()
    rc = lib.cs_create(C.byref(ctx), 0, 1, 256, 256)
    assert rc < 0 and not ctx.value
    assert b"no CPU fallback" in lib.cs_last_error(None)
    from canonswap_b200.engine import CanonSwapError, Engine
    with pytest.raises(CanonSwapError):
        Engine({}, net_hw=(256, 256))


def test_missing_library_fails_loudly(tmp_path):
    """The product path has no fallback: a missing libcanonswap_b200.so (here: CANONSWAP_B200_LIB pointing at a path that does
    not exist, in a fresh interpreter) must raise, naming the file and the build command."""
    import subprocess
    import sys
    env = dict(os.environ, CANONSWAP_B200_LIB=str(tmp_path / "absent.so"))
    code = "from canonswap_b200 import _lib\ntry:\n    _lib.load()\nexcept RuntimeError as e:\n    print('RAISED', e)\n"
    out = subprocess.run([sys.executable, "-c", code], env=env, cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert "RAISED" in out.stdout and "absent.so" in out.stdout and "no CPU / PyTorch fallback" in out.stdout, out.stdout + out.stderr


def test_create_rejects_bad_arguments(lib):
    import ctypes as C
    ctx = C.c_void_p()
    assert lib.cs_create(C.byref(ctx), 0, 0, 256, 256) == -1        # max_batch
    assert lib.cs_create(C.byref(ctx), 0, 1, 200, 256) == -1        # not a multiple of 128
    assert lib.cs_create(None, 0, 1, 256, 256) == -1
    assert lib.cs_frame(None, None, None, None, None, None, 1, 0, None) == -1
    assert lib.cs_launch_count(None) == 0


def test_mirror_modules_keep_reference_state_dict_layout(synth_w):
    from canonswap_b200 import modules, spec
    from canonswap_b200.engine import CanonSwapError
    sw = modules.can_swapper(weights=None, device_id=0)
    mods = {"appearance_feature_extractor": sw.appearance_feature_extractor, "warping_module": sw.warping_module,
            "spade_generator": sw.spade_generator, "transfer": sw.swap_module, "refine": sw.refine_module}
    for name, m in mods.items():
        assert set(m.state_dict().keys()) == set(spec.net_spec(name).keys())
        r = m.load_state_dict(synth_w[name], strict=True)
        assert not r.missing_keys and not r.unexpected_keys
        for k, v in m.state_dict().items():
            assert torch.equal(v, synth_w[name][k])
        assert not m.training
    with pytest.raises(CanonSwapError):
        sw.appearance_feature_extractor(torch.zeros(1, 3, 256, 256))          # CPU tensor -> no fallback
    with pytest.raises(CanonSwapError):
        sw.refine_module.train()
    assert modules.transfer_model_big is modules.transfer_model2


def test_wrapper_host_helpers_match_oracle():
    from canonswap_b200 import modules
    from oracle import canonswap_oracle as O
    sw = modules.can_swapper.__new__(modules.can_swapper)
    sw.device, sw.input_shape = "cpu", (16, 16)
    u8 = np.random.RandomState(0).randint(0, 256, (3, 16, 16, 3)).astype(np.uint8)
    y = sw.prepare_videos([f for f in u8])
    assert torch.equal(y, O.prepare_videos(torch.from_numpy(u8)))
    x = sw.prepare_source(u8[0])
    assert x.shape == (1, 3, 16, 16)
    img = torch.rand(2, 3, 8, 8) * 1.2 - 0.1
    assert np.array_equal(sw.parse_output(img), O.parse_output(img).numpy())


def test_shard_indices_partition():
    from canonswap_b200.pipeline import batches, shard_indices
    for T in (0, 1, 7, 64, 2048):
        for world in (1, 2, 4, 8):
            allv = sorted(i for r in range(world) for i in shard_indices(T, r, world))
            assert allv == list(range(T))
            for r in range(world):
                assert all(i % world == r for i in shard_indices(T, r, world))
    assert batches(list(range(5)), 2) == [[0, 1], [2, 3], [4]]
    assert batches([], 8) == []
    with pytest.raises(ValueError):
        shard_indices(4, 2, 2)


_GLOO_WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["CS_ROOT"])
from canonswap_b200.pipeline import broadcast_identity, run_sharded
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ["CS_PORT"],
                        rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
T, B = 37, 4
g = torch.Generator().manual_seed(3)
frames = torch.randint(0, 256, (T, 4, 4, 3), generator=g, dtype=torch.uint8)
sid = torch.nn.functional.normalize(torch.randn(1, 512, generator=g)) if rank == 0 else None
sid = broadcast_identity(sid)
out = torch.zeros(T, 4, 4, 3, dtype=torch.int64)
def proc(ids):            # stand-in for cs_frame: any per-frame function of (frame, identity)
    for i in ids:
        out[i] = frames[i].long() * 3 + int(sid[0, i % 512].item() * 1e6) % 7
n = run_sharded(proc, T, B, rank, world)
cnt = torch.tensor([n]); dist.all_reduce(cnt)
dist.all_reduce(out)      # test-only reassembly (the product returns frames by per-rank D2H)
if rank == 0:
    np.save(os.environ["CS_OUT"], out.numpy())
    np.save(os.environ["CS_OUT"] + ".sid.npy", sid.numpy())
    assert cnt.item() == T
dist.destroy_process_group()
'''


def test_world_size_2_gloo_sharding_reassembles_exactly(tmp_path):
    """N>1 path on CPU: round-robin shards + one identity broadcast == the single-process result."""
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    out = str(tmp_path / "out.npy")
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", CS_PORT=str(port), CS_OUT=out, CS_ROOT=ROOT)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    for p in procs:
        o, _ = p.communicate(timeout=240)
        assert p.returncode == 0, o.decode()
    got = np.load(out)
    sid = torch.from_numpy(np.load(out + ".sid.npy"))
    g = torch.Generator().manual_seed(3)
    frames = torch.randint(0, 256, (37, 4, 4, 3), generator=g, dtype=torch.uint8)
    sid0 = torch.nn.functional.normalize(torch.randn(1, 512, generator=g))
    assert torch.equal(sid, sid0)                                  # broadcast id equals the root's bit-for-bit
    exp = torch.stack([frames[i].long() * 3 + int(sid0[0, i % 512].item() * 1e6) % 7 for i in range(37)])
    assert np.array_equal(got, exp.numpy())


def test_can_swapper_constructor_behaves_like_the_reference(tmp_path, monkeypatch, synth_w):
    """`can_swapper(inference_cfg)` alone (reference can_swap_e2e.py:44-100): reads models_config, loads
    pretrained_weights/combined_weights.pth when it exists, keeps going without it (as the reference's load_cpk does)."""
    import types
    from canonswap_b200 import modules
    from canonswap_b200.engine import CanonSwapError
    monkeypatch.chdir(tmp_path)
    yaml_path = None
    for root in ("/root/reference", os.path.join(ROOT, "oracle", "_ref")):
        p = os.path.join(root, "src", "config", "models.yaml")
        if os.path.exists(p):
            yaml_path = p
    cfg = types.SimpleNamespace(device_id=0, flag_force_cpu=False, flag_use_half_precision=False, flag_do_torch_compile=False,
                                models_config=yaml_path, input_shape=(256, 256))
    sw = modules.can_swapper(cfg)                                      # no checkpoint in the cwd: zero weights, no error
    assert sw.netArc is None and float(sw.refine_module.state_dict()["resblocks1.0.conv1.weight"].abs().sum()) == 0.0
    os.makedirs("pretrained_weights")
    small = {k: synth_w[k] for k in ("appearance_feature_extractor", "refine")}
    full = dict(synth_w)
    torch.save({**full, **small}, "pretrained_weights/combined_weights.pth")
    sw = modules.can_swapper(cfg)                                      # the reference's two-argument-free construction
    for name, m in (("appearance_feature_extractor", sw.appearance_feature_extractor), ("refine", sw.refine_module),
                    ("transfer", sw.swap_module)):
        for k, v in m.state_dict().items():
            assert torch.equal(v, synth_w[name][k]), (name, k)
    assert sw.device == "cuda:0" and sw.input_shape == (256, 256)
    if yaml_path:                                                      # a config this library does not implement is refused
        import yaml
        bad = yaml.safe_load(open(yaml_path))
        bad["model_params"]["appearance_feature_extractor_params"]["num_resblocks"] = 4
        yaml.safe_dump(bad, open("bad.yaml", "w"))
        cfg.models_config = "bad.yaml"
        with pytest.raises(CanonSwapError):
            modules.can_swapper(cfg)
    cfg.models_config = yaml_path
    cfg.flag_force_cpu = True
    with pytest.raises(CanonSwapError):
        modules.can_swapper(cfg)


_GLOO_V2I_WORKER = r'''
import os, sys, types, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["CS_ROOT"])
from canonswap_b200.pipeline import V2IPipeline, FramePipeline, shard_indices
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ["CS_PORT"],
                        rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
sw = types.SimpleNamespace(device="cpu")
vp = V2IPipeline.__new__(V2IPipeline)                     # host logic only: no engine, no CUDA
vp.sw, vp.batch, vp.net_h, vp.net_w, vp.dev, vp.state = sw, 4, 128, 128, torch.device("cpu"), None
g = torch.Generator().manual_seed(5)
if rank == 0:
    vp.state = {k: torch.randn(*s, generator=g) for k, s in vp._shapes().items()}
    vp.state["swap_can"] = torch.zeros(1)                 # rank-0-only extras are not broadcast
st = vp.broadcast(src=0)
np.save(os.environ["CS_OUT"] + ".%d.npy" % rank, torch.cat([st[k].reshape(-1) for k in V2IPipeline.STATE_KEYS]).numpy())
dist.barrier()
dist.destroy_process_group()
'''


def test_world_size_2_gloo_v2i_state_broadcast(tmp_path):
    """The v2i pipeline's one collective (the per-source state: appearance volume + keypoints / pose, packed into ONE buffer)
    on a world of 2 over gloo: every rank ends with rank 0's state bit for bit, in the declared shapes."""
    import socket
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_V2I_WORKER)
    out = str(tmp_path / "state")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", CS_PORT=str(port), CS_OUT=out, CS_ROOT=ROOT)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    for p in procs:
        o, _ = p.communicate(timeout=240)
        assert p.returncode == 0, o.decode()
    a, b = np.load(out + ".0.npy"), np.load(out + ".1.npy")
    assert a.shape == b.shape and a.size == 32 * 16 * 32 * 32 + 63 + 63 + 9 + 3 + 1
    assert np.array_equal(a, b)
    g = torch.Generator().manual_seed(5)
    from canonswap_b200.pipeline import V2IPipeline
    vp = V2IPipeline.__new__(V2IPipeline)
    vp.net_h = vp.net_w = 128
    exp = torch.cat([torch.randn(*s, generator=g).reshape(-1) for s in vp._shapes().values()]).numpy()
    assert np.array_equal(a, exp)
