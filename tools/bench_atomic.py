"""Supplementary throughput measurement of the `fix eph/atomic` device path (not part of bench.py's contract; the
BASELINE configurations all use `fix eph`).  Device-resident x, v, f (torch tensors), CUDA events on the engine's
stream, W warm-up + K timed steps of post_force + end_of_step.

    python tools/bench_atomic.py --cells 60 --steps 20 --warmup 3 [--loops 2] [--flags 7]

Prints one JSON line: atom-steps/s, ms per step, mean neighbours per atom."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "user-eph_b200"))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=60, help="fcc unit cells per box edge (60 -> 864 000 atoms)")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--loops", type=int, default=2, help="inner loops of the heat-diffusion solve")
    ap.add_argument("--flags", type=int, default=7)
    a = ap.parse_args()
    import torch
    from eph_b200 import atomic as A
    from eph_harness import harness as H
    from eph_b200 import host
    assert torch.cuda.is_available(), "needs a GPU: there is no CPU fallback"
    dev = torch.device("cuda", 0)
    s = H.make_system(a.cells)
    st = torch.cuda.Stream()
    torch.cuda.set_stream(st)
    eng = A.AtomicEngine([0], [0], a.flags, inner_loops=a.loops, stream=st.cuda_stream)
    eng.set_tables_from(host.BetaTables(path=os.path.join(ROOT, "tests", "golden", "Ni_trunc.beta")),
                        A.KappaTables(os.path.join(ROOT, "tests", "golden", "synth1.kappa")))
    eng.set_dt(1e-4)
    t = lambda arr, dt: torch.as_tensor(np.ascontiguousarray(arr), dtype=dt, device=dev)
    keep = [t(s["type"], torch.int32), t(s["mask"], torch.int32), t(s["tag"], torch.int64), t(s["ghost_owner"], torch.int32),
            t(s["offsets"], torch.int64), t(s["neigh"], torch.int32)]
    eng.set_atoms(s["nlocal"], s["nghost"], *keep[:4])
    eng.set_neighbors(*keep[4:])
    eng.init_energy(300.0)
    x, v = t(s["x"], torch.float64), t(s["v"], torch.float64)
    f = torch.zeros((s["nlocal"], 3), dtype=torch.float64, device=dev)

    def step(k):
        eng.post_force(x, v, f, None, k)
        eng.lib.eph_b200_atomic_end_of_step(eng.h, None, None)   # no host read-back inside the timed region

    for k in range(a.warmup):
        step(k + 1)
    beg, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    beg.record(st)
    for k in range(a.steps):
        step(a.warmup + k + 1)
    end.record(st)
    torch.cuda.synchronize()
    ms = beg.elapsed_time(end) / a.steps
    Ee, Te = eng.summary()
    print(json.dumps({"metric": "fix eph/atomic atom-steps/s", "value": s["nlocal"] / (ms * 1e-3), "unit": "atom-steps/s",
                      "ms_per_step": ms, "atoms": s["nlocal"], "ghosts": s["nghost"], "mean_neighbours": float(len(s["neigh"])) / s["nlocal"],
                      "flags": a.flags, "inner_loops": a.loops, "Ee": Ee, "Te": Te, "gpu_launches": eng.launch_count(),
                      "dtype": "f64", "data": "synthetic"}))


if __name__ == "__main__":
    main()
