"""TEST AND BENCHMARK SUPPORT -- not part of the product (user-eph_b200/).

Plays LAMMPS around the engine for the tests and bench.py: `harness` builds synthetic systems (fcc lattices, Maxwell
velocities, periodic ghost images, the full neighbour list, synthetic parameter files; its C++ helper is
libeph_harness.so, built by the Makefile next to it), `parallel` decomposes a box into LAMMPS-style bricks and works out
who owns which ghost.  Nothing under user-eph_b200/ imports this package.
"""
import os
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PKG_DIR)
LIB = os.path.join(PKG_DIR, "libeph_harness.so")


def build(verbose=False):
    subprocess.check_call(["make", "-C", PKG_DIR], stdout=None if verbose else subprocess.DEVNULL)
    if not os.path.exists(LIB):
        raise RuntimeError("build did not produce %s" % LIB)
