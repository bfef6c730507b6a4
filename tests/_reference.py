"""Locating the UNMODIFIED reference package for tests that drive it: the offline install `baseline/_ref` (travels to the GPU
box) or the read-only checkout /root/reference (build container).  Test infrastructure only."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_path():
    for p in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.isdir(os.path.join(p, "specforge")):
            return p
    return None


def import_reference():
    """Puts the reference on sys.path (with the CPU shims the reference's own tests use) and returns its path, or None."""
    p = reference_path()
    if p is None:
        return None
    os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
    if p not in sys.path:
        sys.path.insert(0, p)
    return p
