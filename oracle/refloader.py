"""Load the UNMODIFIED reference ``utmos.select`` from /root/reference (TEST INFRASTRUCTURE ONLY).

This module is part of ``oracle/`` -- the checker, never the product.  Only ``tests/``, the golden-vector
generator (``oracle/make_golden.py``) and nothing on the shipped path may import it.  It only works in
the build container: ``/root/reference`` does not exist on the GPU box.

The reference cannot be imported as-is because ``h5py``, ``truvari`` and ``allel`` are not installed
(utmos/select.py:10,12 and utmos/convert.py:9-12).  They are only needed for IO / logging, not for the
selection arithmetic, so we register inert stub modules (recipe: SURVEY.md Appendix B) and import the
reference source file untouched.
"""
import importlib
import logging
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("UTMOS_REFERENCE_ROOT", "/root/reference")


def reference_available():
    """True when the reference tree is mounted (build container only)."""
    return os.path.exists(os.path.join(REFERENCE_ROOT, "utmos", "select.py"))


def load_reference_select():
    """Return the reference ``utmos.select`` module object, imported unmodified."""
    if not reference_available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_ROOT}")
    if "h5py" not in sys.modules:
        h5 = types.ModuleType("h5py")
        h5.File = type("File", (), {})
        h5.Dataset = type("Dataset", (), {})
        sys.modules["h5py"] = h5
    if "truvari" not in sys.modules:
        tv = types.ModuleType("truvari")

        def setup_logging(debug=False, **_kwargs):
            logging.basicConfig(stream=sys.stderr, level=logging.DEBUG if debug else logging.INFO)
        tv.setup_logging = setup_logging
        sys.modules["truvari"] = tv
    if "allel" not in sys.modules:
        sys.modules["allel"] = types.ModuleType("allel")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    return importlib.import_module("utmos.select")
