#!/usr/bin/env python
"""Recipe for oracle/_ref/: the UNMODIFIED reference modules of the hot path, copied byte for byte from
/root/reference/libs at build time so that they travel to the GPU box (which has no /root/reference) and can be
timed there as the CPU arm of bench.py (`--impl reference`, `cpu_baseline.kind = "reference"`).

oracle/_ref/ is git-ignored (reference sources never enter this repository's history) but not gpurun-ignored.
Test infrastructure only: nothing under esn-ofdm-mimo_b200/ imports it.

    python oracle/build_ref.py          # also run by __graft_entry__.build() when /root/reference exists
"""
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIBS = "/root/reference/libs"
OUT = os.path.join(HERE, "_ref")
FILES = ("pyESN.py", "helper_mimo_esn_generic.py", "HelpFunc.py")


def build_ref():
    """Returns the output directory, or None when the reference is not present (the GPU box)."""
    if not os.path.isdir(REF_LIBS):
        return OUT if os.path.exists(os.path.join(OUT, FILES[0])) else None
    os.makedirs(OUT, exist_ok=True)
    manifest = {}
    for f in FILES:
        src = os.path.join(REF_LIBS, f)
        shutil.copyfile(src, os.path.join(OUT, f))
        with open(src, "rb") as fh:
            manifest[f] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as fh:
        json.dump({"source": REF_LIBS, "sha256": manifest}, fh, indent=1)
    return OUT


def load_ref():
    """Import the copied reference modules under private names (they must not shadow the drop-in `pyESN`)."""
    import importlib.util
    path = os.path.join(OUT, "pyESN.py")
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location("_reference_pyESN", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print("oracle/_ref:", build_ref())
