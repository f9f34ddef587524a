"""General (non-uniform) FDM path timing: python tools/fdm_bench_general.py [n]"""
import sys, time, numpy as np
sys.path.insert(0, "user-eph_b200"); sys.path.insert(0, "."); sys.path.insert(0, "tests")
from eph_b200 import lib, host
import test_gpu_parity as T
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
eng = lib.Engine([0], flags=7)
eng.set_tables_from(host.BetaTables(path="tests/golden/Ni_trunc.beta"))
L = 56.32 * n / 128.0
rng = np.random.default_rng(1)
kap = 0.01248 * (0.9 + 0.1 * rng.random(n ** 3))
eng.set_grid(n, n, n, [0, L, 0, L, 0, L], 300.0, 1.0, 3.5e-6, kap)
eng.set_dt(1e-4)
T._solve_only(eng); T._solve_only(eng)
eng.synchronize()
eng.set_profiling(True)
for _ in range(3):
    T._solve_only(eng)
kt = eng.kernel_times()["fdm_substep"]
ms = kt[0] / kt[1]
print("general path: %.4f ms per sub-step, %.1f Gcell-updates/s, %.0f GB/s algorithmic (60 B/cell), substeps %d" % (ms, n ** 3 / ms / 1e6, 60 * n ** 3 / ms / 1e6, eng.last_substeps()))
