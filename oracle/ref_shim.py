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


class _RefFinder:
    """Import `models.x` / `utils.x` / `engines.x` from oracle/_ref/<pkg>/<x>.pyc.bin (byte-compiled reference modules)."""

    @staticmethod
    def find_spec(name, path=None, target=None):
        import importlib.machinery as M
        import importlib.util as U
        parts = name.split(".")
        if parts[0] not in ("models", "utils", "engines") or len(parts) > 2:
            return None
        if len(parts) == 1:
            d = os.path.join(REF, parts[0])
            if not os.path.isdir(d):
                return None
            spec = M.ModuleSpec(name, None, is_package=True)
            spec.submodule_search_locations = [d]
            return spec
        f = os.path.join(REF, parts[0], parts[1] + ".pyc.bin")
        if not os.path.isfile(f):
            return None
        return U.spec_from_file_location(name, f, loader=M.SourcelessFileLoader(name, f))


if available() and KIND == "pyc":
    sys.meta_path.insert(0, _RefFinder)
if available():
    _stub("imageio")
    mp = _stub("matplotlib")
    mp.pyplot = _stub("matplotlib.pyplot")
    _stub("lpips", LPIPS=_LPIPS)
    _stub("sacrebleu", dataset=None)          # engines/trainer.py:8 `from sacrebleu import dataset` (unused there)
    _stub("configargparse")
    if REF not in sys.path:
        sys.path.insert(0, REF)
