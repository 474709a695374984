"""TEST INFRASTRUCTURE ONLY.

Makes the UNMODIFIED reference importable in the build container:
``/root/reference`` (read-only OQuPy 0.5.0) on top of ``oracle/tn_shim`` (the
restated third-party ``tensornetwork`` slice + import stubs).  /root/reference
does not exist on the GPU box, so nothing in ``-m gpu`` tests, ``smoke()`` or
``bench.py`` may call this; it is used by ``tests/golden/make_golden.py`` and by
CPU tests that are skipped when the reference tree is absent.
"""
import os
import sys

REFERENCE_ROOT = "/root/reference"
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tn_shim")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "oqupy"))


def load_reference():
    """Return the imported reference ``oqupy`` module (shimmed third parties)."""
    if not reference_available():
        raise RuntimeError("reference tree not present (expected in the build "
                           "container only)")
    for p in (_SHIM, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import oqupy  # pylint: disable=import-outside-toplevel
    return oqupy
