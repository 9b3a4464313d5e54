"""Command line driver: the role of the reference's SpatialCL (SpatialCL/main.cpp:23-69) on the B200 library, plus the
volume CSV of outputTimeCource (SpatialSBML/outputFunction.cpp:30-205: header lines, then one row `x, y[, z], <species...>`
per volume cell).

  python -m spatial_sbml_b200.cli model.xml [--dims 101 101 1] [--end 1001] [--dt 0.01] [--out DIR] [--every N]

Defaults are SpatialCL's: a 101 x 101 grid and run(1001, 0.01), i.e. oneStep(t, dt) for t = 0 .. int(end / dt)
(spatialsimulator.cpp:1667-1703).  Needs a CUDA device: there is no CPU fallback.
"""
import argparse
import os
import sys
import time

import numpy as np

from . import SpatialSimulator


def spatial_species(sim):
    names = []
    for line in sim.getSetupText("summary").splitlines():
        p = line.split()
        if p and p[0] == "var" and p[3] == "0" and p[7] == "1":
            names.append(p[1])
    return names


def write_volume_csv(path, sim, names, model_name, end_time, dt, sim_time):
    nz, ny, nx = sim.localShape()
    X, Y, Z = 2 * nx - 1, 2 * ny - 1, 2 * nz - 1
    x = sim.getX().reshape(Z, Y, X)[::2, ::2, ::2]
    y = sim.getY().reshape(Z, Y, X)[::2, ::2, ::2]
    cols = [x.ravel(), y.ravel()]
    head = "x, y"
    dz = None
    if nz > 1:
        z = sim.getZ().reshape(Z, Y, X)[::2, ::2, ::2]
        cols.append(z.ravel())
        head = "x, y, z"
        dz = z[1, 0, 0] - z[0, 0, 0]
    fields = [sim.getVariableCompact(n).ravel() for n in names]
    with open(path, "w") as f:
        f.write("# results of %s.xml simulation\n" % model_name)
        f.write("# simulation time = %g, dt = %g, current time = %g\n" % (end_time, dt, sim_time))
        f.write("# dx = %g" % (x[0, 0, 1] - x[0, 0, 0]))
        if ny > 1:
            f.write(", dy = %g" % (y[0, 1, 0] - y[0, 0, 0]))
        if dz is not None:
            f.write(", dz = %g" % dz)
        f.write("\n\n# " + head + "".join(", " + n for n in names) + "\n")
        np.savetxt(f, np.column_stack(cols + fields), fmt="%.17g", delimiter=", ")


def main(argv=None):
    ap = argparse.ArgumentParser(prog="spatial_sbml_b200.cli", description=__doc__.split("\n\n")[0])
    ap.add_argument("model")
    ap.add_argument("--dims", type=int, nargs="+", default=[101, 101, 1])
    ap.add_argument("--end", type=float, default=1001.0)
    ap.add_argument("--dt", type=float, default=0.01)
    ap.add_argument("--out", default=None, help="directory for <step>.csv volume files (none: no files)")
    ap.add_argument("--every", type=int, default=0, help="write a file every N steps (0: only the last state)")
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args(argv)
    dims = (a.dims + [1, 1])[:3]
    t0 = time.time()
    sim = SpatialSimulator(device=a.device)
    sim.initFromFile(a.model, *dims)
    names = spatial_species(sim)
    model_name = os.path.splitext(os.path.basename(a.model))[0]
    if a.out:
        os.makedirs(a.out, exist_ok=True)
    last = int(a.end / a.dt)          # steps t = 0 .. last, as in run()
    total = last + 1
    # steps are issued in chunks that end where something has to be printed or written
    marks = {total}
    if last >= 10:
        marks.update((last // 10) * pct + 1 for pct in range(11) if (last // 10) * pct + 1 <= total)
    if a.out and a.every:
        marks.update(range(a.every, total + 1, a.every))
    t = 0
    percent = 0
    for stop in sorted(marks):
        if stop > t:
            sim.runSteps(t, a.dt, stop - t)
            t = stop
        while last >= 10 and percent <= 10 and t - 1 >= (last // 10) * percent:
            print("%d%% finished" % (percent * 10), flush=True)   # spatialsimulator.cpp:1696-1699
            percent += 1
        if a.out and a.every and t % a.every == 0 and t < total:
            write_volume_csv(os.path.join(a.out, "%d.csv" % t), sim, names, model_name, a.end, a.dt, (t - 1) * a.dt)
    if a.out:
        write_volume_csv(os.path.join(a.out, "%d.csv" % t), sim, names, model_name, a.end, a.dt, (t - 1) * a.dt)
    print("time: %.3f" % (time.time() - t0), file=sys.stderr)
    sim.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
