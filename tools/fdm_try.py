import sys, numpy as np
sys.path.insert(0, "user-eph_b200"); sys.path.insert(0, "."); sys.path.insert(0, "tests")
from eph_b200 import lib, host
import test_gpu_parity as T
eng = lib.Engine([0], flags=7)
eng.set_tables_from(host.BetaTables(path="tests/golden/Ni_trunc.beta"))
n = 64
eng.set_grid(n, n, n, [0, 10, 0, 10, 0, 10], 300.0, 1.0, 3.5e-6, 0.1248)
eng.set_dt(1e-6)
T._solve_only(eng)
print("T mean", eng.mean_T(), "substeps", eng.last_substeps())
