#!/usr/bin/env python
"""Byte-compile the reference's own files for the render path into oracle/_ref/ (git-ignored, NOT gpurun-ignored).

TEST / BASELINE INFRASTRUCTURE ONLY.  The reference is pure Python and /root/reference does not exist on the GPU box, so
`bench.py --impl reference` could otherwise only time the op-for-op port (oracle/torch_port.py).  This recipe is the
Python analogue of compiling a C reference into oracle/_ref/*.so: the UNMODIFIED sources are compiled where they lie
(py_compile, no source text is written anywhere in this repo) and only the resulting bytecode binaries (*.pyc.bin, loaded by
importlib's SourcelessFileLoader under the same interpreter version) land in oracle/_ref/.  `oracle/ref_shim.py` installs an
importer for them when /root/reference is absent.  Run by __graft_entry__.build() whenever /root/reference is present.
"""
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("NSOS_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
# the hot path (models/nerf_net.py:132 and everything it calls), the two loss modules + metrics the trainer/eval drop-ins are
# checked against, and the stock training step
FILES = ["models/nerf_net.py", "models/nerf_mlp.py", "models/sampler.py", "models/embedder.py", "models/renderer.py",
         "utils/error.py", "utils/image.py", "utils/ssim.py", "utils/ray.py", "utils/misc.py", "utils/get_metrics.py",
         "engines/lr.py"]


def build(verbose=True) -> bool:
    if not os.path.isdir(os.path.join(REF, "models")):
        return False
    for rel in FILES:
        src = os.path.join(REF, rel)
        dst = os.path.join(OUT, rel[:-3] + ".pyc.bin")          # models/nerf_net.pyc.bin (gpurun's snapshot skips *.pyc)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if os.path.exists(dst) and os.path.getmtime(dst) >= os.path.getmtime(src):
            continue
        py_compile.compile(src, cfile=dst, dfile=rel, doraise=True)
    with open(os.path.join(OUT, "PYTHON_TAG"), "w") as f:
        f.write(sys.implementation.cache_tag + "\n")
    if verbose:
        print(f"oracle/_ref: {len(FILES)} reference modules byte-compiled from {REF} ({sys.implementation.cache_tag})")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
