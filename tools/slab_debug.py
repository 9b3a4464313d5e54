"""Debug aid: 1 device vs slabs on one device, where do they first differ."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import spatial_sbml_b200 as sb
from spatial_sbml_b200 import synthetic

def run(label, model, dims, dt, steps, nslabs, species=("species_1", "species_2")):
    a = sb.SpatialSimulator(); a.initFromString(model, *dims)
    b = sb.SpatialSimulator(devices=[0] * nslabs); b.initFromString(model, *dims)
    for t in range(steps):
        a.oneStep(t, dt); b.oneStep(t, dt)
        for sp in species:
            x, y = a.getVariableCompact(sp), b.getVariableCompact(sp)
            if not np.array_equal(x, y):
                bad = np.argwhere(x != y)
                print(label, "step", t + 1, sp, "differs:", len(bad), "cells; planes", sorted(set(bad[:, 0].tolist())), "rows", sorted(set(bad[:, 1].tolist()))[:12],
                      "cols", bad[:, 2].min(), bad[:, 2].max(), "max abs", np.abs(x - y).max(), flush=True)
                a.close(); b.close()
                return
    print(label, "identical after", steps, "steps", flush=True)
    a.close(); b.close()

dims = (18, 14, 10)
bc_all = {("species_1", "Xmin"): 0.3, ("species_1", "Xmax"): -0.1, ("species_1", "Ymin"): 0.2, ("species_1", "Ymax"): 0.05,
          ("species_1", "Zmin"): 0.15, ("species_1", "Zmax"): -0.07, ("species_2", "Xmin"): 0.02, ("species_2", "Ymax"): 0.04,
          ("species_2", "Zmax"): (0.12, "Dirichlet"), ("species_2", "Zmin"): (0.09, "Dirichlet")}
run("all", synthetic.turing(*dims, bc_values=bc_all), dims, 0.02, 4, 2)
run("neumann xy only", synthetic.turing(*dims, bc_values={k: v for k, v in bc_all.items() if k[1][0] in "XY"}), dims, 0.02, 4, 2)
run("neumann z only", synthetic.turing(*dims, bc_values={k: v for k, v in bc_all.items() if k[1][0] == "Z" and not isinstance(v, tuple)}), dims, 0.02, 4, 2)
run("dirichlet z only", synthetic.turing(*dims, bc_values={k: v for k, v in bc_all.items() if isinstance(v, tuple)}), dims, 0.02, 4, 2)
run("neumann x only", synthetic.turing(*dims, bc_values={("species_1", "Xmin"): 0.3}), dims, 0.02, 4, 2)
run("2d neumann x only", synthetic.turing(18, 14, bc_values={("species_1", "Xmin"): 0.3}), (18, 14, 1), 0.02, 4, 2)
