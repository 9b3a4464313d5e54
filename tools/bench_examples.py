"""The reference's example models (BASELINE configs 1-4) and small synthetic grids: µs per oneStep and cell-updates/s on
the GPU, next to the reference's own serial CPU step (oracle/_ref/ref_driver) timed in the same run on the same box.
usage: python tools/bench_examples.py [gpu steps] [--json out.json]"""
import json
import os
import subprocess
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import torch
import spatial_sbml_b200 as sb

steps = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 400
out_json = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
DRIVER = os.path.join(REPO, "oracle", "_ref", "ref_driver")
rows = []
for case in ("turing_tiny", "turing_tiny_cl", "turing_blob", "brusselator_300", "fish_300", "advection", "transport", "syn_turing_3d"):
    g = np.load(os.path.join(REPO, "tests", "golden", case + ".npz"))
    nx, ny, nz = [int(v) for v in g["dims"]]
    dt = float(g["dt"])
    model = os.path.join(REPO, "tests", "golden", "models", case + ".xml")
    sim = sb.SpatialSimulator()
    sim.initFromFile(model, nx, ny, nz)
    def timed(first):
        torch.cuda.synchronize()
        t0 = time.time()
        sim.runSteps(first, dt, steps)
        torch.cuda.synchronize()
        return 1e6 * (time.time() - t0) / steps
    sim.runSteps(0, dt, 20)
    interp_us = timed(20)          # first steps of a simulator: the interpreter kernel
    specialised = sim.specializeNow()  # what a long run switches to by itself after 256 steps (background build)
    sim.runSteps(20 + steps, dt, 20)
    l0 = sim.getDeviceLaunchCount()
    gpu_us = timed(40 + steps)
    cells = sim.getNumDomainCells()
    launches = (sim.getDeviceLaunchCount() - l0) / steps
    sim.close()
    cpu = None
    if os.path.exists(DRIVER):
        csteps = max(3, min(200, int(2.0e6 * 3.0 / max(cells, 1))))  # about 3 s of CPU work
        r = subprocess.run([DRIVER, "time", model, str(nx), str(ny), str(nz), repr(dt), "2", str(csteps)], capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if line:
            cpu = json.loads(line[-1])
    cpu_us = 1e6 * cpu["seconds"] / cpu["steps"] if cpu else float("nan")
    rows.append({"case": case, "dims": [nx, ny, nz], "cells": int(cells), "gpu_us_per_step": gpu_us, "gpu_us_per_step_interpreter": interp_us,
                 "specialised": bool(specialised), "launches_per_step": launches,
                 "gpu_cell_updates_per_s": cells / (gpu_us * 1e-6), "cpu_us_per_step": cpu_us,
                 "cpu_cell_updates_per_s": cells / (cpu_us * 1e-6) if cpu else None, "speedup": cpu_us / gpu_us if cpu else None})
    print("%-16s %4dx%4dx%3d  GPU %8.1f us/step (interpreter kernel %6.1f; %4.1f launches) %9.2f M cu/s | reference CPU (1 thread) %10.1f us/step %6.2f M cu/s | x%.0f" % (
        case, nx, ny, nz, gpu_us, interp_us, launches, cells / gpu_us, cpu_us, cells / cpu_us if cpu else 0.0, cpu_us / gpu_us if cpu else 0.0), flush=True)
if out_json:
    json.dump({"host_cores": os.cpu_count(), "rows": rows}, open(out_json, "w"), indent=1)
