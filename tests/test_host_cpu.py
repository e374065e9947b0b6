"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol include/nerfsos.h declares,
the flat parameter layout matches the reference state_dict, the drop-in keeps the reference's
checkpoint keys / seeded initialisation, and the product path refuses to run without CUDA."""
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden

import nerfsos_b200  # noqa: F401
from nerfsos_b200 import _lib
from nerfsos_b200.models.nerf_net import NeRFNet


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "nerfsos.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(nsos_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.nsos_abi_version() == 1


def test_param_layout_matches_reference_state_dict():
    sd = load_golden("flower_weights")["sd"]
    L = _lib.lib()
    d = _lib.NetDesc(8, 256, 4, 10, 4, 1, 1, 2, 1)
    n = L.nsos_param_count(d)
    keys = [k for k in sd if k.startswith("nerf.mlp.")]
    assert n == sum(sd[k].size for k in keys) == 637062
    import ctypes as C
    cap = 64
    offs = (C.c_int64 * cap)(); rows = (C.c_int32 * cap)(); cols = (C.c_int32 * cap)()
    cnt = L.nsos_param_layout(d, offs, rows, cols, cap)
    assert cnt == len(keys)
    off = 0
    for i, k in enumerate(keys):                      # state_dict order == flat order
        shape = sd[k].shape
        assert offs[i] == off, k
        assert rows[i] == shape[0] and cols[i] == (shape[1] if len(shape) == 2 else 1), k
        off += sd[k].size
    # config[0] net and the packed-image sizes the tcgen05 path reports
    d1 = _lib.NetDesc(4, 64, 4, 10, 4, 1, 1, 2, 1)
    g = load_golden("cfg1_d4w64_eval")["sd"]
    assert L.nsos_param_count(d1) == sum(v.size for k, v in g.items() if k.startswith("nerf.mlp."))
    assert L.nsos_packed_bytes(d, _lib.MODE_TC_EXACT) > L.nsos_packed_bytes(d, _lib.MODE_TC_FAST) > 0
    assert L.nsos_packed_bytes(d, _lib.MODE_SIMT) == 0
    # invalid: D == skip+1 (the reference cannot run it either) and W not covered by the tcgen05 path
    assert L.nsos_param_count(_lib.NetDesc(5, 256, 4, 10, 4, 1, 1, 2, 1)) < 0
    assert L.nsos_packed_bytes(_lib.NetDesc(8, 100, 4, 10, 4, 1, 1, 2, 1), _lib.MODE_TC_EXACT) == 0


def test_dropin_keeps_reference_keys_and_seeded_init():
    g = load_golden("cfg1_d4w64_eval")
    torch.manual_seed(0)
    net = NeRFNet(netdepth=4, netwidth=64, netdepth_fine=4, netwidth_fine=64, N_samples=64, N_importance=0,
                  use_semantics=True, sem_with_coord=True)
    sd = net.state_dict()
    assert list(sd) == list(g["sd"])
    for k in sd:                                       # same construction order => bit-identical default init
        assert np.array_equal(sd[k].numpy(), g["sd"][k]), k
    net2 = NeRFNet(N_samples=64, N_importance=128, use_semantics=True, sem_with_coord=True, sem_dim=2)
    fl = load_golden("flower_weights")["sd"]
    assert list(net2.state_dict()) == list(fl)
    net2.load_state_dict({k: torch.from_numpy(v) for k, v in fl.items()}, strict=True)
    # stage-1 checkpoints have no semantic head: strict=False load leaves exactly those 8 keys missing
    stage1 = {k: torch.from_numpy(v) for k, v in fl.items() if "semantic_linear" not in k}
    r = net2.load_state_dict(stage1, strict=False)
    assert len(r.missing_keys) == 8 and all("semantic_linear" in k for k in r.missing_keys) and not r.unexpected_keys
    assert net2.nerf is not net2.nerf_fine
    assert NeRFNet(netdepth=2, netwidth=64, N_samples=8, N_importance=0).nerf_fine is not None
    n0 = NeRFNet(netdepth=2, netwidth=64, N_samples=8, N_importance=0)
    assert n0.nerf_fine is n0.nerf                      # nerf_net.py:49
    assert net2.render_kwargs_test["perturb"] == 0. and net2.render_kwargs_test["raw_noise_std"] == 0.


def test_no_cpu_fallback():
    net = NeRFNet(netdepth=2, netwidth=64, netdepth_fine=2, netwidth_fine=64, N_samples=8, N_importance=0)
    with pytest.raises(_lib.NsosError):
        net(torch.zeros(2, 4, 3), (1.0, 2.0))
    with pytest.raises(_lib.NsosError):
        net.nerf(torch.zeros(4, 3), viewdirs=torch.zeros(4, 3))


def test_product_path_never_imports_oracle():
    """The oracle is test infrastructure: nothing under nerf-sos_b200/ may import or call it."""
    pkg = os.path.join(ROOT, "nerf-sos_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "nerf_oracle" not in src, f


def test_sincos_cw_constants_accuracy():
    """The positional encoding's sine/cosine (csrc/common.cuh: sincos_cw) restated in numpy with the constants parsed from
    the CUDA source: <= 1.6 ulp / 8e-8 absolute against float64 on the encoder's argument range |2^k x| <= 8192."""
    import re
    src = open(os.path.join(ROOT, "nerf-sos_b200", "csrc", "common.cuh")).read()
    body = src[src.index("void sincos_cw("):src.index("// embedder.py:34-48, column c")]
    num = r"(-?[0-9.]+(?:e[-+]?[0-9]+)?)f"
    two_over_pi = float(re.search(r"__fmaf_rn\(x, " + num, body).group(1))
    cw = [float(m) for m in re.findall(r"__fmaf_rn\(q, " + num, body)]
    ps = [float(m) for m in re.findall(r"ps = __fmaf_rn\((?:ps|" + num[:-1] + r"f), r2, " + num, body)[0] if m] + \
         [float(m[-1]) for m in re.findall(r"ps = __fmaf_rn\((ps), r2, " + num, body)]
    pc = [float(m) for m in re.findall(r"pc = __fmaf_rn\((?:pc|" + num[:-1] + r"f), r2, " + num, body)[0] if m] + \
         [float(m[-1]) for m in re.findall(r"pc = __fmaf_rn\((pc), r2, " + num, body)]
    assert len(cw) == 3 and len(ps) == 4 and len(pc) == 4, (cw, ps, pc)
    f32 = np.float32

    def fma(a, b, c):
        return (np.float64(a) * np.float64(b) + np.float64(c)).astype(f32)

    rng = np.random.default_rng(0)
    x = rng.uniform(-16, 16, 400_000).astype(f32)
    for k in (0, 3, 6, 9):
        a = (x * f32(2 ** k)).astype(f32)
        q = np.rint(fma(a, f32(two_over_pi), f32(0))).astype(f32)
        i = q.astype(np.int64)
        r = a
        for c in cw:
            r = fma(q, f32(c), r)
        r2 = (r * r).astype(f32)
        s = f32(ps[0])
        for c in ps[1:]:
            s = fma(s, r2, f32(c))
        sn = fma(s, (r2 * r).astype(f32), r)
        c_ = f32(pc[0])
        for c in pc[1:]:
            c_ = fma(c_, r2, f32(c))
        cs = fma(c_, r2, f32(1))
        swap = (i & 1) == 1
        S, C = np.where(swap, cs, sn), np.where(swap, sn, cs)
        S = np.where((i & 2) == 2, -S, S)
        C = np.where(((i + 1) & 2) == 2, -C, C)
        ts, tc = np.sin(np.float64(a)), np.cos(np.float64(a))
        assert np.abs(S - ts).max() < 8e-8 and np.abs(C - tc).max() < 8e-8, k
        big = np.abs(ts) > 0.1
        assert (np.abs(S - ts)[big] / np.spacing(np.abs(ts[big]).astype(f32))).max() < 1.6, k


def test_bench_has_no_undefined_names():
    """bench.py's GPU branches cannot run here; at least every name they load must be bound somewhere in the function,
    the module or builtins (a copy-paste slip in the train line once only showed up on the 8-GPU box)."""
    import ast
    import builtins
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    module_names = {n.id for n in ast.walk(tree) if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Store)}
    module_names |= {n.name for n in ast.walk(tree) if isinstance(n, (ast.FunctionDef, ast.ClassDef))}
    for n in ast.walk(tree):
        if isinstance(n, (ast.Import, ast.ImportFrom)):
            module_names |= {(a.asname or a.name).split(".")[0] for a in n.names}
    for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)]:
        bound = {a.arg for a in fn.args.args + fn.args.kwonlyargs} | module_names
        for n in ast.walk(fn):
            if isinstance(n, ast.ExceptHandler) and n.name:
                bound.add(n.name)
            if isinstance(n, (ast.Lambda, ast.FunctionDef)):
                bound |= {a.arg for a in n.args.args}
            if isinstance(n, ast.comprehension):
                bound |= {t.id for t in ast.walk(n.target) if isinstance(t, ast.Name)}
        loads = {n.id for n in ast.walk(fn) if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load)}
        missing = sorted(x for x in loads if x not in bound and not hasattr(builtins, x))
        assert not missing, (fn.name, missing)


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the unmodified reference on the host cores -- from /root/reference here, from the
    byte-compiled oracle/_ref on the GPU box, the oracle port only if neither imports) runs without a GPU and prints exactly
    one JSON line with the contract's keys, on the FULL 4096-ray configuration."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "rays/s" and j["value"] > 0 and j["higher_is_better"] is True
    assert j["cpu_baseline"]["kind"] in ("reference", "port") and j["cpu_baseline"]["cores"] >= 1 and j["e2e"]["h2d_bytes_per_step"] == 0
    assert j["config"]["rays_per_step"] == 4096 and j["steps"] == 1
    assert {"metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config"} <= set(j)


def test_bench_line_builders_run_on_cpu():
    """Every --workload arm's line-building code executes here, so that a NameError cannot take down an 8-GPU run."""
    import argparse
    import json
    import bench
    for mode in ("exact", "fast"):
        args = argparse.Namespace(steps=10, warmup=3, mode=mode, all_params=False)
        m = dict(total_ms=30.0, e2e_s=0.04, e2e_dropin_s=0.06, n_total=8 * 4096, n_local=4096, t_wall=0.1, clk={"sm_mhz": 1900.0},
                 h2d=98304, d2h=278528)
        train = dict(ms_per_step=50.0, rays_per_s=1e6, rays_per_gpu=32768, steps=5, warmup=3, trainable_params=82436,
                     recipe="--fix_backbone (semantic heads)", collectives_per_step=4, allgather_bytes_per_rank=1, allreduce_bytes=2,
                     h2d_bytes_per_step=3, d2h_bytes_per_step=4, final_loss=0.1, clocks=None)
        for image in (False, True):
            line = bench.eval_line(m, args, 8, image, None if image else train)
            json.dumps(line)
            assert line["n_gpus"] == 8 and line["roofline"]["frac"] > 0 and ("train" in line) == (not image)
        for rec in ("--fix_backbone (semantic heads)", "all parameters"):
            tl = bench.train_line(dict(train, recipe=rec), args, 8)
            json.dumps(tl)
            assert tl["collectives"]["collectives_per_step"] == 4 and tl["roofline"]["frac"] > 0
