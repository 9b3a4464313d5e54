"""Debug aid: oneStepHost against oneStep at the benchmark size, step after step (which step / rows first differ)."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import torch
import spatial_sbml_b200 as sb
from spatial_sbml_b200 import synthetic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 24
pre = int(sys.argv[3]) if len(sys.argv) > 3 else 100
model = synthetic.turing(n, n)
names = ["species_1", "species_2"]
a = sb.SpatialSimulator(); a.initFromString(model, n, n, 1)
b = sb.SpatialSimulator(); b.initFromString(model, n, n, 1)
a.runSteps(0, 0.07, pre); b.runSteps(0, 0.07, pre)
hb = [torch.empty((1, n, n), dtype=torch.float64).pin_memory().numpy() for _ in names]
ho = [torch.empty((1, n, n), dtype=torch.float64).pin_memory().numpy() for _ in names]
for x, nm in zip(hb, names):
    b.getVariableCompact(nm, out=x)
for t in range(steps):
    a.oneStep(pre + t, 0.07)
    b.oneStepHost(pre + t, 0.07, dict(zip(names, hb)), dict(zip(names, ho)), 0)
    hb, ho = ho, hb
    for x, nm in zip(hb, names):
        ref = a.getVariableCompact(nm)
        fin = np.isfinite(x).all()
        same = np.array_equal(x, ref)
        if not (fin and same):
            bad = np.argwhere(~((x == ref) | (np.isnan(x) & np.isnan(ref))))
            rows = np.unique(bad[:, 1])
            print("step", t, nm, "finite", fin, "ref finite", np.isfinite(ref).all(), "differs at", len(bad), "cells; rows", rows[:10], "...", rows[-5:], "cols", bad[:, 2].min(), bad[:, 2].max(), flush=True)
            sys.exit(1)
    print("step", t, "ok", float(np.abs(hb[0]).max()), flush=True)
print("ALL OK")
