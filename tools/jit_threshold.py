"""Per-step time of the interpreter kernel and of the model-specialised build over grid sizes (GPU box).
usage: python tools/jit_threshold.py [sizes...]   → one JSON line per size"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import spatial_sbml_b200 as sb
from spatial_sbml_b200 import synthetic


def run(n, mode, dim3=False):
    os.environ["SB2_JIT"] = mode
    dims = (n, n, n) if dim3 else (n, n, 1)
    sim = sb.SpatialSimulator()
    t0 = time.time()
    sim.initFromString(synthetic.turing(n, n, n if dim3 else None), *dims)
    init = time.time() - t0
    cells = n * n * (n if dim3 else 1)
    steps = max(20, min(2000, int(2e9 / cells / 4)))
    dt = 0.02 if dim3 else 0.07
    sim.runSteps(0, dt, 20)
    torch.cuda.synchronize()
    t0 = time.time()
    sim.runSteps(20, dt, steps)
    torch.cuda.synchronize()
    return (time.time() - t0) / steps * 1e6, init, sim.isSpecialized()


if __name__ == "__main__":
    sizes = [a for a in sys.argv[1:]] or ["256", "512", "1024", "2048", "4096"]
    for s in sizes:
        dim3 = s.endswith("c")
        n = int(s.rstrip("c"))
        a, ia, sa = run(n, "0", dim3)
        b, ib, sbb = run(n, "1", dim3)
        print(json.dumps({"grid": "%d^%d" % (n, 3 if dim3 else 2), "interpreter_us": round(a, 2), "specialised_us": round(b, 2),
                          "speedup": round(a / b, 3), "init_s": [round(ia, 2), round(ib, 2)], "specialised": [sa, sbb]}), flush=True)
