"""ms per step of the synthetic Brusselator on an nx x ny grid (what a slab of a strongly scaled grid looks like to one GPU).
usage: python tools/strong_probe.py nx ny"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, spatial_sbml_b200 as sb
from spatial_sbml_b200 import synthetic
nx, ny = int(sys.argv[1]), int(sys.argv[2])
sim = sb.SpatialSimulator(); sim.initFromString(synthetic.brusselator(nx, ny), nx, ny, 1)
sim.runSteps(0, 0.1, 10)
best = 1e9
for r in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sim.runSteps(10, 0.1, 40); torch.cuda.synchronize()
    best = min(best, (time.perf_counter() - t0) / 40 * 1e3)
print("%dx%d %.4f ms/step specialised %s" % (nx, ny, best, sim.isSpecialized()))
