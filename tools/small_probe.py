"""One small example model, specialised kernel, `steps` steps through runSteps (graph replay): prints us per step.  Under
`ncu --metrics gpu__time_duration.sum` the launch list shows what the kernels themselves take.  usage: small_probe.py case [steps]"""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import torch
import spatial_sbml_b200 as sb
case = sys.argv[1] if len(sys.argv) > 1 else "brusselator_300"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 400
g = np.load(os.path.join(REPO, "tests", "golden", case + ".npz"))
nx, ny, nz = [int(v) for v in g["dims"]]
dt = float(g["dt"])
sim = sb.SpatialSimulator()
sim.initFromFile(os.path.join(REPO, "tests", "golden", "models", case + ".xml"), nx, ny, nz)
sim.runSteps(0, dt, 20)
sim.specializeNow()
sim.runSteps(20, dt, 20)
first = 40
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sim.runSteps(first, dt, steps)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    us = 1e6 * (time.perf_counter() - t0) / steps
    first += steps
    print("%s %dx%dx%d: %.2f us/step (the call returned after %.2f us/step: host time to queue the steps)" % (case, nx, ny, nz, us, 1e6 * (t1 - t0) / steps), flush=True)
sim.close()
