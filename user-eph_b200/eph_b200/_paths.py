"""Locations of the in-tree shared libraries and the commands that build them."""
import os
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(os.path.dirname(PKG_DIR))

_LIBS = {
    "engine": os.path.join(PKG_DIR, "libeph_b200.so"),
    "fix": os.path.join(PKG_DIR, "libeph_b200_fix.so"),
    "atomic_fix": os.path.join(PKG_DIR, "libeph_b200_atomic_fix.so"),
}


def lib_path(name):
    # EPH_B200_ENGINE_LIB: development aid for A/B timing of two builds of the engine on the same box
    if name == "engine" and os.environ.get("EPH_B200_ENGINE_LIB"):
        return os.environ["EPH_B200_ENGINE_LIB"]
    return _LIBS[name]


def build_all(verbose=False):
    """Compile the CUDA library for sm_100a (nvcc cross-compiles without a GPU) and the host libraries."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", os.path.join(PKG_DIR, "..", "csrc")], stdout=out)
    subprocess.check_call(["make", "-C", os.path.join(PKG_DIR, "..", "fix")], stdout=out)
    for p in _LIBS.values():
        if not os.path.exists(p):
            raise RuntimeError("build did not produce %s" % p)
