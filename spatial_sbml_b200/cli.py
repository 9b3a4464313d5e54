"""Command line driver: the role of the reference's SpatialCL (SpatialCL/main.cpp:23-69) on the B200 library, plus the
text output of outputTimeCource (SpatialSBML/outputFunction.cpp:30-205): the volume CSV (header lines, then one row
`x, y[, z], <species...>` per volume cell) and, for models with species on membranes, the membrane CSV (one row per point of
the doubled grid: `x, y[, z]` and `mFlag, value` per membrane species; mFlag 0 / 1 / 2 = off the membrane / membrane /
pseudo-membrane point), written to <out>/volume/<n>.csv and <out>/membrane/<n>.csv like the reference's result tree.

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


def membrane_species(sim):
    """(species id, compartment id) of the species that live on a membrane (variableInfo::inVol false)"""
    membranes, out = set(), []
    lines = [l.split() for l in sim.getSetupText("summary").splitlines()]
    for p in lines:
        if p and p[0] == "geo" and p[p.index("isVol") + 1] == "0":
            membranes.add(p[1])
    for p in lines:
        if p and p[0] == "var" and p[3] == "0" and "geo" in p and p[p.index("geo") + 1] in membranes and p[5] == "1":
            out.append((p[1], p[p.index("geo") + 1]))
    return out


def write_membrane_csv(path, sim, mem, model_name, end_time, dt, sim_time):
    """outputFunction.cpp:65-98 (header) and :133-152 / :175-198 (rows over EVERY point of the doubled grid)"""
    nz, ny, nx = sim.localShape()
    X, Y, Z = 2 * nx - 1, 2 * ny - 1, 2 * nz - 1
    x, y = sim.getX()[:X * Y * Z], sim.getY()[:X * Y * Z]
    cols = [x, y]
    head = "x, y"
    if nz > 1:
        cols.append(sim.getZ()[:X * Y * Z])
        head = "x, y, z"
    for sid, comp in mem:
        cols.append(np.asarray(sim.getGeometry(comp), dtype=np.float64)[:X * Y * Z])
        cols.append(np.asarray(sim.getVariable(sid))[:X * Y * Z])
        sim.releaseVariable(sid)
    table = np.column_stack(cols)
    fmt = ["%.17g"] * (2 + (nz > 1)) + ["%d", "%.17g"] * len(mem)
    with open(path, "w") as f:
        f.write("# results of %s.xml simulation\n" % model_name)
        f.write("# simulation time = %g, dt = %g, current time = %g\n" % (end_time, dt, sim_time))
        f.write("# about mFlag: whether sid is not in the membrane domain(0), in the membrane domain(1), pseudo membrane domain(2)\n")
        f.write("# dx = %g" % (x[2] - x[0]))
        if ny > 1:
            f.write(", dy = %g" % (y[2 * X] - y[0]))
        if nz > 1:
            f.write(", dz = %g" % (cols[2][2 * X * Y] - cols[2][0]))
        f.write("\n\n# " + head + "".join(", mFlag, " + sid for sid, _ in mem) + "\n")
        for r in range(Y * Z):  # a blank line after every row of points (3-D: except after the last row of a plane)
            np.savetxt(f, table[r * X:(r + 1) * X], fmt=fmt, delimiter=", ")
            if nz == 1 or (r % Y) != Y - 1:
                f.write("\n")


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
    mem = membrane_species(sim)
    names = [n for n in spatial_species(sim) if n not in set(sid for sid, _ in mem)]  # volume species (variableInfo::inVol)
    model_name = os.path.splitext(os.path.basename(a.model))[0]
    if a.out:
        os.makedirs(a.out, exist_ok=True)
        if mem:
            os.makedirs(os.path.join(a.out, "volume"), exist_ok=True)
            os.makedirs(os.path.join(a.out, "membrane"), exist_ok=True)

    def write_all(step):
        # models without membrane species keep the flat <out>/<n>.csv layout
        vol = os.path.join(a.out, "volume", "%d.csv" % step) if mem else os.path.join(a.out, "%d.csv" % step)
        write_volume_csv(vol, sim, names, model_name, a.end, a.dt, (step - 1) * a.dt)
        if mem:
            write_membrane_csv(os.path.join(a.out, "membrane", "%d.csv" % step), sim, mem, model_name, a.end, a.dt, (step - 1) * a.dt)

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
            write_all(t)
    if a.out:
        write_all(t)
    print("time: %.3f" % (time.time() - t0), file=sys.stderr)
    sim.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
