"""Import shim for the read-only reference tree (build container only).  TEST INFRASTRUCTURE ONLY.

Inserts stub modules for the reference's missing, non-arithmetic imports and puts
/root/reference on sys.path so `models.*` / `utils.*` import unmodified.  Used by
`oracle/make_golden.py`, by `tests/test_oracle_vs_reference.py` and by bench.py's reference arm.  Where /root/reference is
absent (the GPU box) it falls back to oracle/_ref/: byte-compiled (.pyc) copies of the unmodified reference modules produced
by oracle/build_ref.py in the build container.
"""
import os
import sys
import types

REF = os.environ.get("NSOS_REFERENCE", "/root/reference")
KIND = "source"
if not os.path.isdir(os.path.join(REF, "models")):
    # GPU box: the byte-compiled copies of the unmodified reference modules made by oracle/build_ref.py (same interpreter)
    _pyc = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
    _tag = os.path.join(_pyc, "PYTHON_TAG")
    if os.path.isfile(_tag) and open(_tag).read().strip() == sys.implementation.cache_tag:
        REF, KIND = _pyc, "pyc"


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
