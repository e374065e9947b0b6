"""Launch time of the general weight-gradient kernel (dW[256, 320] += dY[P, 256]^T . [X[P, 256] | gamma[P, 64]]) on P = 1.5 M
points (one 8192-ray chunk of the fine net), NSOS_WGRAD_V1=1 (row-owning threads) and the
cp.async-staged version with 8 / 16 fill warps (NSOS_WGRAD_W8=1), requests by the fill threads (NSOS_WGRAD_NOLOADER=1) or by four loader warps (default).  Also checks that the two agree.  usage: python tools/time_wgrad.py [P]"""
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
    g = torch.Generator(device=dev).manual_seed(0)
    dy = torch.randn(P, 256, device=dev, generator=g) * 1e-3
    x = torch.relu(torch.randn(P, 256, device=dev, generator=g))
    e = torch.randn(P, 64, device=dev, generator=g)
    scr = torch.empty(L.nsos_selftest_wgrad_scratch_bytes(), dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    res = {}
    variants = (("v1", ("NSOS_WGRAD_V1",)), ("v2 8w self", ("NSOS_WGRAD_W8", "NSOS_WGRAD_NOLOADER")), ("v2 16w self", ("NSOS_WGRAD_NOLOADER",)),
                ("v2 8w+4ld", ("NSOS_WGRAD_W8",)), ("v2 16w+4ld", ()))
    for tag, envs, aux_w in [(n, e, a) for a in (0, 63) for n, e in variants]:
        tag = tag + ("+aux" if aux_w else "")
        for k in ("NSOS_WGRAD_V1", "NSOS_WGRAD_W8", "NSOS_WGRAD_NOLOADER"):
            os.environ.pop(k, None)
        for k in envs:
            os.environ[k] = "1"
        dw = torch.zeros(256, 320, device=dev)
        db = torch.zeros(256, device=dev)
        ts = []
        for i in range(6):
            flush.fill_(0)
            dw.zero_()
            db.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(L.nsos_selftest_wgrad(_lib.ptr(dy), 256, 256, _lib.ptr(x), 256, 64, _lib.ptr(e if aux_w else None), 64, aux_w, 0,
                                             _lib.ptr(dw), 320, _lib.ptr(db), P, _lib.ptr(scr), scr.numel(), None), "wgrad")
            e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1))
        ms = sum(ts) / len(ts)
        gb = P * (256 * 4 * 2 + (256 if aux_w else 0)) / 1e9          # algorithmic bytes: dY once, X once (+ gamma)
        print(f"{tag:16s} {ms:7.3f} ms  {gb / ms:6.2f} TB/s algorithmic  (P = {P})")
        res[tag] = (dw.clone(), db.clone())
    for tag in res:
        ref = res["v1+aux" if tag.endswith("+aux") else "v1"]
        d = (ref[0] - res[tag][0]).abs().max().item() / ref[0].abs().max().item()
        dbias = (ref[1] - res[tag][1]).abs().max().item() / max(ref[1].abs().max().item(), 1e-30)
        print(f"v1 vs {tag}: dW rel diff {d:.2e}, db rel diff {dbias:.2e}")


if __name__ == "__main__":
    main()
