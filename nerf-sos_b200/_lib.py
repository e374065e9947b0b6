"""ctypes binding of libnerfsos.so (C ABI: include/nerfsos.h) + the in-tree nvcc build recipe."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "lib", "libnerfsos.so")
if os.environ.get("NSOS_LIB"):      # A/B testing of kernel variants: load another build of the same library
    LIB_PATH = os.path.abspath(os.environ["NSOS_LIB"])
SOURCES = ["api.cu", "simt_gemm.cu", "simt_render.cu", "tc_render.cu", "tc_wgrad.cu", "corr_loss.cu", "optim.cu"]

MODE_SIMT, MODE_TC_EXACT, MODE_TC_FAST = 0, 1, 2
MODES = {"simt": MODE_SIMT, "exact": MODE_TC_EXACT, "fast": MODE_TC_FAST}


class NsosError(RuntimeError):
    pass


class NetDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("D", "W", "skip", "multires", "multires_views", "use_viewdirs",
                                         "use_semantics", "sem_dim", "sem_with_coord")]


class RenderCfg(C.Structure):
    _fields_ = [("coarse", NetDesc), ("fine", NetDesc), ("n_samples", C.c_int32), ("n_importance", C.c_int32),
                ("perturb", C.c_float), ("raw_noise_std", C.c_float), ("white_bkgd", C.c_int32), ("mode", C.c_int32)]


class Randoms(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("t_rand", "noise0", "u", "noise1", "z_samples")]


class RenderOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("maps", "weights0", "weights", "raw0", "raw", "z_vals0", "z_vals",
                                          "z_samples", "inds", "h_last0", "s_hid0", "h_last", "s_hid", "enc0", "enc", "status")]


class AdamTensor(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p), ("n", C.c_int64)]


class LossShard(C.Structure):
    _fields_ = [("q0", C.c_int32), ("nq", C.c_int32), ("B_total", C.c_int32), ("phase", C.c_int32), ("sums", C.c_void_p)]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into nerf-sos_b200/lib/libnerfsos.so (nvcc cross-compiles
    without a GPU).  Rebuilds only when a source is newer than the library."""
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(HERE, "..", "include", "nerfsos.h"))
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
           "-shared", "-o", LIB_PATH] + srcs
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise NsosError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    """Load libnerfsos.so.  Fails loudly if it has not been built -- there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NsosError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(needs nvcc).  nerfsos_b200 has no CPU or PyTorch fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, sz, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t, C.c_uint64
    P = C.POINTER
    sig = {
        "nsos_abi_version": (C.c_int, []),
        "nsos_last_error": (C.c_char_p, []),
        "nsos_param_count": (i64, [P(NetDesc)]),
        "nsos_param_layout": (C.c_int, [P(NetDesc), P(i64), P(i32), P(i32), C.c_int]),
        "nsos_packed_bytes": (sz, [P(NetDesc), C.c_int]),
        "nsos_pack_weights": (C.c_int, [P(NetDesc), vp, vp, C.c_int, vp]),
        "nsos_render_workspace_bytes": (sz, [P(RenderCfg), i64]),
        "nsos_render_fwd": (C.c_int, [P(RenderCfg), vp, vp, vp, vp, vp, vp, vp, vp, P(Randoms), u64, P(RenderOut), vp, sz, i64, vp]),
        "nsos_render_bwd_workspace_bytes": (sz, [P(RenderCfg), i64, C.c_int]),
        "nsos_render_bwd": (C.c_int, [P(RenderCfg), vp, vp, vp, vp, vp, vp, vp, vp, P(Randoms), u64, vp, vp, vp, C.c_int, P(RenderOut), vp, sz, i64, vp]),
        "nsos_invert_cdf": (C.c_int, [vp, vp, vp, vp, vp, i64, i32, i32, vp]),
        "nsos_mlp_workspace_bytes": (sz, [P(NetDesc), i64]),
        "nsos_mlp_query": (C.c_int, [P(NetDesc), vp, vp, vp, vp, vp, sz, i64, vp]),
        "nsos_geo_corr_workspace_bytes": (sz, [i32, i32, i32]),
        "nsos_geo_corr_loss": (C.c_int, [vp, vp, vp, vp, vp, vp, i32, i32, i32, vp, sz, vp]),
        "nsos_app_corr_workspace_bytes": (sz, [i32, i32, i32, i32]),
        "nsos_app_corr_loss": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp, sz, vp]),
        "nsos_geo_corr_loss_sharded": (C.c_int, [vp, vp, vp, vp, vp, vp, i32, i32, i32, P(LossShard), vp, sz, vp]),
        "nsos_app_corr_loss_sharded": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, P(LossShard), vp, sz, vp]),
        "nsos_adam_multi": (C.c_int, [P(AdamTensor), i32, C.c_float, C.c_float, C.c_float, C.c_float, i64, vp]),
        "nsos_selftest_rowgemm": (C.c_int, [vp, i64, i32, vp, i64, i64, vp, i64, i32, vp, i64, vp, C.c_int, C.c_int, i64, vp, sz, vp]),
        "nsos_selftest_wgrad": (C.c_int, [vp, i64, i32, vp, i64, i32, vp, i64, i32, i32, vp, i64, vp, i64, vp, C.c_size_t, vp]),
        "nsos_selftest_wgrad_scratch_bytes": (C.c_size_t, []),
        "nsos_mlp_query_dir": (C.c_int, [C.POINTER(NetDesc), vp, vp, C.POINTER(C.c_float), vp, i32, i64, vp]),
        "nsos_selftest_umma": (C.c_int, [vp, vp, vp, i32, i32, C.c_int, C.c_int, vp, sz, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)          # AttributeError here == the .so does not export what nerfsos.h declares
        fn.restype, fn.argtypes = res, args
    if L.nsos_abi_version() != 1:
        raise NsosError("libnerfsos.so ABI version mismatch")
    _lib = L
    return L


EXPORTS = ["nsos_abi_version", "nsos_last_error", "nsos_param_count", "nsos_param_layout", "nsos_packed_bytes",
           "nsos_pack_weights", "nsos_render_workspace_bytes", "nsos_render_fwd", "nsos_render_bwd_workspace_bytes",
           "nsos_render_bwd", "nsos_invert_cdf", "nsos_mlp_workspace_bytes", "nsos_mlp_query",
           "nsos_geo_corr_workspace_bytes", "nsos_geo_corr_loss", "nsos_app_corr_workspace_bytes", "nsos_app_corr_loss",
           "nsos_geo_corr_loss_sharded", "nsos_app_corr_loss_sharded", "nsos_adam_multi", "nsos_selftest_umma", "nsos_selftest_rowgemm",
           "nsos_selftest_wgrad", "nsos_selftest_wgrad_scratch_bytes", "nsos_mlp_query_dir"]


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().nsos_last_error().decode(errors="replace")
        raise NsosError(f"{what} failed (status {rc}): {msg}")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def cur_stream(device):
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
