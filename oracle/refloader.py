"""Load the UNMODIFIED reference ``utmos.select`` (TEST INFRASTRUCTURE ONLY).

This module is part of ``oracle/`` -- the checker, never the product.  Only ``tests/``, the golden-vector
generators (``oracle/make_golden*.py``) and ``bench.py``'s CPU-baseline legs may import it; nothing on the shipped
path does.

Where the reference comes from: ``oracle/_ref/`` when it exists -- the reference package installed UNMODIFIED by
``build_ref()`` below (``pip install --no-index --no-deps --target oracle/_ref`` of a scratch copy of
``/root/reference``; build output, git-ignored, travels to the GPU box like the built ``.so`` files) -- else
``/root/reference`` itself (build container only).  No reference source is ever copied into the repository.

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
HERE = os.path.dirname(os.path.abspath(__file__))
INSTALLED_ROOT = os.path.join(HERE, "_ref")


def build_ref(force=False):
    """Install the unmodified reference package into oracle/_ref (only where /root/reference is mounted).

    The reference's setup.py writes into its source tree and /root/reference is read-only, so pip runs on a
    scratch copy.  Dependencies (h5py, scikit-allel, truvari: no wheels here) are not resolved: --no-deps.
    Returns the install directory, or None when the reference tree is absent (GPU box: the prebuilt copy is used)."""
    import shutil
    import subprocess
    import tempfile
    if not os.path.exists(os.path.join(REFERENCE_ROOT, "utmos", "select.py")):
        return INSTALLED_ROOT if os.path.exists(os.path.join(INSTALLED_ROOT, "utmos", "select.py")) else None
    if not force and os.path.exists(os.path.join(INSTALLED_ROOT, "utmos", "select.py")):
        return INSTALLED_ROOT
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "src")
        shutil.copytree(REFERENCE_ROOT, src, ignore=shutil.ignore_patterns(".git", "repo_utils"))
        shutil.rmtree(INSTALLED_ROOT, ignore_errors=True)
        subprocess.check_call([sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation",
                               "--no-deps", "--target", INSTALLED_ROOT, src],
                              stdout=subprocess.DEVNULL, stderr=subprocess.STDOUT)
    return INSTALLED_ROOT


def reference_root():
    """Directory to put on sys.path: the installed copy if present, else the mounted tree, else None."""
    for root in (INSTALLED_ROOT, REFERENCE_ROOT):
        if os.path.exists(os.path.join(root, "utmos", "select.py")):
            return root
    return None


def reference_available():
    """True when the reference can be imported (installed copy in oracle/_ref or the mounted tree)."""
    return reference_root() is not None


def load_reference_select():
    """Return the reference ``utmos.select`` module object, imported unmodified."""
    root = reference_root()
    if root is None:
        raise RuntimeError(f"reference neither installed in {INSTALLED_ROOT} nor mounted at {REFERENCE_ROOT}")
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
    if root not in sys.path:
        sys.path.insert(0, root)
    return importlib.import_module("utmos.select")
