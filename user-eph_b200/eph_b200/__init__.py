"""eph_b200 -- Python side of the B200-native `fix eph` hot path.

Thin ctypes bindings over the C ABI (include/eph_b200.h) and over the host-side
C++ of the product (table construction, grid files, the FixEPHB200 class driven
through the LAMMPS stand-in).  PyTorch is used only for device memory, streams
and torch.distributed plumbing.  What plays LAMMPS for the tests and bench.py
(synthetic systems, brick decomposition) lives outside the product, in
eph_harness/ at the repository root.
"""
from ._paths import PKG_DIR, REPO_ROOT, build_all, lib_path  # noqa: F401
