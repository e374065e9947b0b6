"""Import shim for the read-only reference tree (build container only).  TEST INFRASTRUCTURE ONLY.

Inserts stub modules for the reference's missing, non-arithmetic imports and puts
/root/reference on sys.path so `models.*` / `utils.*` import unmodified.  Used by
`oracle/make_golden.py` and by the optional `tests/test_oracle_vs_reference.py` (skipped when
/root/reference is absent, e.g. on the GPU box).
"""
import os
import sys
import types

REF = os.environ.get("NSOS_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "models"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _LPIPS:
    def __init__(self, **kw):
        pass

    def __call__(self, *a, **k):
        raise RuntimeError("lpips is stubbed")


if available():
    _stub("imageio")
    mp = _stub("matplotlib")
    mp.pyplot = _stub("matplotlib.pyplot")
    _stub("lpips", LPIPS=_LPIPS)
    _stub("sacrebleu")
    _stub("configargparse")
    if REF not in sys.path:
        sys.path.insert(0, REF)
