"""`.dmp` concentration dumps of the reference's front ends (SpatialSBML/datahelper.hh:78-114; SpatialUI's
importConcentration / exportConcentration, spatialmainwidget.cpp:205-257): a text file "maxX maxY" followed by maxX lines of
maxY numbers, written with the default iostream formatting (six significant digits), indexed data[x][y].

    helper = DataHelper.from_simulator(sim, "species_1")   # what exportConcentration collects: getVariableAt(x, y)
    helper.write("s1.dmp")
    DataHelper("s1.dmp").apply(sim, "species_1")           # what importConcentration does: a dose at every cell of the domain
"""
import math

import numpy as np


def _fmt(v):
    """operator<<(double) with the default precision of 6 (what datahelper.hh:104 writes)"""
    if v != v:
        return "nan" if math.copysign(1.0, v) > 0 else "-nan"
    if v in (float("inf"), float("-inf")):
        return "inf" if v > 0 else "-inf"
    return "%g" % v


class DataHelper(object):
    def __init__(self, file_name=None, rows=0, cols=0, value=0.0):
        self.data = np.full((rows, cols), float(value), dtype=np.float64)
        self.valid = file_name is None
        if file_name is not None:
            self.read(file_name)

    # datahelper.hh:78-95
    def read(self, file_name):
        try:
            with open(file_name, "rb") as f:
                tok = f.read().split()
        except OSError:
            self.valid = False
            return self
        nx, ny = int(tok[0]), int(tok[1])
        vals = np.zeros(nx * ny, dtype=np.float64)
        have = min(nx * ny, len(tok) - 2)  # a short file leaves the rest at zero, like a failed stream extraction
        vals[:have] = [float(t) for t in tok[2:2 + have]]
        self.data = vals.reshape(nx, ny)
        self.valid = True
        return self

    # datahelper.hh:97-114
    def write(self, file_name):
        nx, ny = self.data.shape
        with open(file_name, "wb") as f:
            f.write(("%d %d\n" % (nx, ny)).encode())
            for x in range(nx):
                f.write(("".join(_fmt(v) + " " for v in self.data[x]) + "\n").encode())

    def is_valid(self):
        return self.valid

    # operator()(double x, double y), datahelper.hh:118-142: floor and clamp to the array
    def _pos(self, x, y):
        nx, ny = self.data.shape
        xp = min(max(int(math.floor(x)), 0), nx - 1)
        yp = min(max(int(math.floor(y)), 0), ny - 1)
        return xp, yp

    def __call__(self, x, y):
        return float(self.data[self._pos(x, y)])

    def set(self, x, y, value):
        self.data[self._pos(x, y)] = value

    @classmethod
    def from_simulator(cls, sim, species):
        """exportConcentration (spatialmainwidget.cpp:233-257): helper(x, y) = the species' value at cell (x, y)"""
        nx, ny = sim.getXDim(), sim.getYDim()
        h = cls(rows=nx, cols=ny)
        field = sim.getVariableCompact(species)
        h.data[:, :] = field[0].T  # compact fields are [z][y][x]
        return h

    def apply(self, sim, species):
        """importConcentration (spatialmainwidget.cpp:205-230 with SimulationThread::applyDose, simulationthread.cpp:160-185):
        every cell of the species' compartment takes the dump's value at that cell; cells outside the compartment are skipped."""
        if not self.valid:
            return False
        nx, ny = sim.getXDim(), sim.getYDim()
        comp = sim.getSpeciesCompartment(species)
        geo = sim.getGeometry(comp) if comp else None
        if geo is None:
            return False
        inside = np.asarray(geo).reshape(-1, 2 * ny - 1, 2 * nx - 1)[0, ::2, ::2] != 0
        field = sim.getVariableCompact(species)
        xs = np.minimum(np.arange(nx), self.data.shape[0] - 1)
        ys = np.minimum(np.arange(ny), self.data.shape[1] - 1)
        dose = self.data[np.ix_(xs, ys)].T
        field[0][inside] = dose[inside]
        sim.setVariableCompact(species, field)
        return True
