"""Launch time of the row GEMM of the all-parameter backward (C[P,256] = (A[P,256] . B[256,256]) masked by M[P,256] > 0) on
P = 1.5 M points (one 8192-ray chunk of the fine net): default (row-owning threads read their mask rows under the MMAs) and
NSOS_RG_MASK_BITS=1 (mask rows read whole and turned into bits with ballots).  Checks 4096 rows against fp64.  usage: python tools/time_rowgemm.py [P]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402


def main():
    import nerfsos_b200  # noqa: F401
    from nerfsos_b200 import _lib
    L = _lib.lib()
    dev = torch.device("cuda", 0)
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 8192 * 192
    K = N = 256
    g = torch.Generator(device=dev).manual_seed(0)
    a = torch.randn(P, K, device=dev, generator=g) * 1e-3
    w = torch.randn(N, K, device=dev, generator=g) * 0.1          # stored [N, K]: B(k, n) = w[n, k]
    mask = torch.randn(P, N, device=dev, generator=g)
    c = torch.empty(P, N, device=dev)
    scratch = torch.zeros(2 * K * N * 2 + 4096, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = torch.randint(0, P, (4096,), device=dev, generator=g)
    ref = torch.where(mask[rows] > 0, a[rows].double() @ w.double().t(), torch.zeros((), dtype=torch.float64, device=dev))
    for tag, env, use_mask, acc in (("mask rows", None, True, 0), ("mask bits", "NSOS_RG_MASK_BITS", True, 0), ("no mask", None, False, 0),
                                    ("no mask, accumulate", None, False, 1)):
        os.environ.pop("NSOS_RG_MASK_BITS", None)
        if env:
            os.environ[env] = "1"
        ts = []
        for i in range(6):
            flush.fill_(0)
            c.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(L.nsos_selftest_rowgemm(_lib.ptr(a), K, K, _lib.ptr(w), 1, K, _lib.ptr(c), N, N, _lib.ptr(mask if use_mask else None), N,
                                               None, 0, acc, P, _lib.ptr(scratch), scratch.numel(), None), "rowgemm")
            e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1))
        ms = sum(ts) / len(ts)
        gb = P * 4 * (K + N + (N if use_mask else 0) + (N if acc else 0)) / 1e9
        err = ((c[rows].double() - ref).abs().max() / ref.abs().max()).item() if use_mask else float("nan")
        print(f"{tag:20s} {ms:7.3f} ms  {gb / ms:5.2f} TB/s algorithmic  {2e-12 * P * K * N / (ms * 1e-3):6.1f} TFLOP/s  max err / max {err:.1e}")


if __name__ == "__main__":
    main()
