"""FDM-only driver for profiling: python tools/fdm_only.py [n] [solves]"""
import sys, numpy as np
sys.path.insert(0, "user-eph_b200"); sys.path.insert(0, "."); sys.path.insert(0, "tests")
from eph_b200 import lib, host
import test_gpu_parity as T
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
solves = int(sys.argv[2]) if len(sys.argv) > 2 else 2
eng = lib.Engine([0], flags=7)
eng.set_tables_from(host.BetaTables(path="tests/golden/Ni_trunc.beta"))
L = 56.32 * n / 128.0
eng.set_grid(n, n, n, [0, L, 0, L, 0, L], 300.0, 1.0, 3.5e-6, 0.01248)
eng.set_dt(1e-4)
for _ in range(solves):
    T._solve_only(eng)
print("T mean", eng.mean_T(), "substeps", eng.last_substeps())
