"""Debug driver: one golden case, a few steps, with the device layer's trace / debug words switched on.
usage: python tools/dbg_rd.py [case] [steps]"""
import os
import sys
import faulthandler
faulthandler.enable()
faulthandler.dump_traceback_later(40, exit=True)
os.environ.setdefault("SB2_TRACE", "1")
os.environ.setdefault("SB2_DEBUG", "1")
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import spatial_sbml_b200 as sb

case = sys.argv[1] if len(sys.argv) > 1 else "turing_tiny"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
g = np.load(os.path.join(REPO, "tests", "golden", case + ".npz"))
nx, ny, nz = [int(v) for v in g["dims"]]
dt = float(g["dt"])
sim = sb.SpatialSimulator()
print("init", case, nx, ny, nz, flush=True)
sim.initFromFile(os.path.join(REPO, "tests", "golden", "models", case + ".xml"), nx, ny, nz)
print("initialised", flush=True)
for t in range(steps):
    sim.oneStep(t, dt)
    print("step", t, "done", flush=True)
for sp in g["species"]:
    got = sim.getVariableCompact(str(sp))
    key = "step/%d/%s" % (steps, sp)
    if key in g.files:
        ref = g[key]
        print(sp, "rel Linf", np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-300), flush=True)
    else:
        print(sp, "finite", np.isfinite(got).all(), "range", got.min(), got.max(), flush=True)
sim.close()
print("ok", flush=True)
