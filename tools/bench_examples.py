"""ms per oneStep of the example models (BASELINE configs 1-4) on the GPU: python tools/bench_examples.py [steps]"""
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import torch
import spatial_sbml_b200 as sb

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 400
for case in ("turing_tiny", "turing_tiny_cl", "brusselator_300", "fish_300", "advection", "transport", "syn_turing_3d"):
    g = np.load(os.path.join(REPO, "tests", "golden", case + ".npz"))
    nx, ny, nz = [int(v) for v in g["dims"]]
    dt = float(g["dt"])
    sim = sb.SpatialSimulator()
    sim.initFromFile(os.path.join(REPO, "tests", "golden", "models", case + ".xml"), nx, ny, nz)
    sim.runSteps(0, dt, 20)
    torch.cuda.synchronize()
    l0 = sim.getDeviceLaunchCount()
    t0 = time.time()
    sim.runSteps(20, dt, steps)
    sim.getVariableCompact(str(g["species"][0]))  # forces completion
    t1 = time.time()
    cells = sim.getNumDomainCells()
    print("%-16s %4dx%4dx%3d  %7.1f us/step  %5.1f launches/step  %8.2f M cell-updates/s" % (
        case, nx, ny, nz, 1e6 * (t1 - t0) / steps, (sim.getDeviceLaunchCount() - l0) / steps, cells * steps / (t1 - t0) / 1e6), flush=True)
    sim.close()
