"""TEST INFRASTRUCTURE — loader for the compiled reference under oracle/_ref/.

`import_reference()` returns the reference's own `MinkowskiEngine` package (its unmodified
Python files on top of its own CPU sources compiled by oracle/build_ref.py), or raises
`ReferenceUnavailable`.  Nothing here reads /root/reference at run time.
"""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


class ReferenceUnavailable(RuntimeError):
    pass


def available():
    return os.path.isfile(os.path.join(REF_DIR, "MinkowskiEngineBackend", "_C.so")) and \
        os.path.isdir(os.path.join(REF_DIR, "MinkowskiEngine"))


_cached = None


def import_reference():
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise ReferenceUnavailable(
            "oracle/_ref is not built: run `python oracle/build_ref.py` where /root/reference exists")
    import torch  # noqa: F401  (the extension links against libtorch)
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    # the reference package imports these optional modules at import time in some files
    sys.modules.setdefault("open3d", types.ModuleType("open3d"))
    try:
        _cached = importlib.import_module("MinkowskiEngine")
    except Exception as e:  # pragma: no cover
        raise ReferenceUnavailable(f"importing the compiled reference failed: {e!r}") from e
    return _cached
