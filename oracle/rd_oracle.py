"""ORACLE / TEST INFRASTRUCTURE — numpy restatement of the reference's explicit reaction-diffusion step.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg may import this module; the product
(spatial_sbml_b200/) never does.

What is restated, in the reference's own evaluation order (all file:line relative to /root/reference/SpatialSBML):
  * the RK4 stage loop of performStep            spatialsimulator.cpp:1350-1492
  * reversePolishRK, the per-cell RPN evaluator   calcPDE.cpp:224-451  (stack discipline, operator semantics,
                                                  reactant -= / product += accumulation)
  * calcDiffusion (uniform and field-valued coefficients)   calcPDE.cpp:453-562
  * updateSpecies (RK4 combine, k := 0)           spatialsimulator.cpp:1518-1544
  * calcBoundary (Neumann flux, Dirichlet value)  calcPDE.cpp:871-1034, incl. the flux carried in k0 between steps
  * cipCSLR (CIP-CSLR advection; uniform and field-valued velocities: face-averaged velocity, divergence correction)  calcPDE.cpp:564-869
  * calcMemTransport (volume species on either side of a membrane, species bound to / released from the membrane itself,
    incl. the rpStack[1] rate quirk)              calcPDE.cpp:1036-1482, call sites spatialsimulator.cpp:1392-1450
  * calcMemDiffusion (species ON a membrane: finite volumes over the Voronoi neighbours)  calcPDE.cpp:1484-1581
  * reactions between species on a membrane (reversePolishRK over the membrane's points) and their update
  * updateAssignmentRules (ordered rules on parameters and species, re-evaluated after every step)  spatialsimulator.cpp:1566-1581
  * the pseudo-membrane fill that ends updateAssignmentRules  spatialsimulator.cpp:1609-1663
Field-valued boundary values, and membrane species as operands of volume reactions or rules, raise NotImplementedError.

Pinned: tests/test_oracle_cpu.py checks this restatement BIT FOR BIT against the fields the unmodified reference
produced (tests/golden/*.npz): every golden case (38), every stored step.

Every array of a volume species is COMPACT (nz, ny, nx): volume cells only; compact cell (i,j,k) is the reference's
doubled-grid point (2i,2j,2k) (spatialsimulator.cpp:354-358), so the reference's index +-2 is +-1 here.  Advected species and
species on membranes keep the reference's whole doubled grid.
"""
import xml.etree.ElementTree as ET

import numpy as np

CORE = "{http://www.sbml.org/sbml/level3/version1/core}"
SPATIAL = "{http://www.sbml.org/sbml/level3/version1/spatial/version1}"

# libSBML ASTNodeType_t values as they appear in the dumped streams ("O <type>")
AST_PLUS, AST_MINUS, AST_TIMES, AST_DIVIDE, AST_POWER = 43, 45, 42, 47, 94
AST_FUNCTION_ABS, AST_FUNCTION_ARCCOS, AST_FUNCTION_ARCCOSH = 269, 270, 271
AST_FUNCTION_ARCCSC, AST_FUNCTION_ARCSIN, AST_FUNCTION_ARCTAN = 274, 278, 280
AST_FUNCTION_COS, AST_FUNCTION_COSH, AST_FUNCTION_CSC, AST_FUNCTION_EXP = 283, 284, 287, 290
AST_FUNCTION_LN, AST_FUNCTION_LOG, AST_FUNCTION_POWER, AST_FUNCTION_ROOT = 293, 294, 296, 297
AST_FUNCTION_SIN, AST_FUNCTION_SINH, AST_FUNCTION_TAN, AST_FUNCTION_TANH = 300, 301, 302, 303
AST_LOGICAL_AND, AST_LOGICAL_NOT, AST_LOGICAL_OR, AST_LOGICAL_XOR = 304, 305, 306, 307
AST_RELATIONAL_EQ, AST_RELATIONAL_GEQ, AST_RELATIONAL_GT = 308, 309, 310
AST_RELATIONAL_LEQ, AST_RELATIONAL_LT, AST_RELATIONAL_NEQ = 311, 312, 313

import math


def _libm(fn, nargs=1):
    """The C library's function applied element by element (numpy's own vectorised exp / pow / tanh ... may differ from libm
    in the last bit, and the reference calls libm); domain errors give NaN / inf like the C function does."""
    def safe(*a):
        try:
            return fn(*a)
        except (ValueError, ZeroDivisionError):
            return float("nan")
        except OverflowError:
            return float("inf")
    uf = np.frompyfunc(safe, nargs, 1)
    return lambda *arrs: uf(*arrs).astype(np.float64)


UNARY = {
    AST_FUNCTION_ABS: np.fabs, AST_FUNCTION_ARCCOS: _libm(math.acos), AST_FUNCTION_ARCCOSH: _libm(math.acosh),
    AST_FUNCTION_ARCCSC: lambda x: _libm(math.asin)(1.0 / x), AST_FUNCTION_ARCSIN: _libm(math.asin), AST_FUNCTION_ARCTAN: _libm(math.atan),
    AST_FUNCTION_COS: _libm(math.cos), AST_FUNCTION_COSH: _libm(math.cosh), AST_FUNCTION_CSC: lambda x: 1.0 / _libm(math.sin)(x),
    AST_FUNCTION_EXP: _libm(math.exp), AST_FUNCTION_LN: _libm(math.log),
    AST_FUNCTION_ROOT: np.sqrt,  # sqrt whatever the degree (calcPDE.cpp:355); correctly rounded everywhere
    AST_FUNCTION_SIN: _libm(math.sin), AST_FUNCTION_SINH: _libm(math.sinh), AST_FUNCTION_TAN: _libm(math.tan), AST_FUNCTION_TANH: _libm(math.tanh),
    AST_LOGICAL_NOT: lambda x: np.where(x.astype(np.int64) == 0, 1.0, 0.0),  # calcPDE.cpp:385-388
}


def _as_int(x):
    with np.errstate(invalid="ignore"):
        return np.nan_to_num(x, nan=0.0, posinf=0.0, neginf=0.0).astype(np.int64)


BINARY = {
    AST_PLUS: lambda a, b: a + b, AST_MINUS: lambda a, b: a - b, AST_TIMES: lambda a, b: a * b, AST_DIVIDE: lambda a, b: a / b,
    AST_POWER: _libm(math.pow, 2), AST_FUNCTION_POWER: _libm(math.pow, 2),
    AST_FUNCTION_LOG: lambda a, b: _libm(math.log)(b) / _libm(math.log)(a),  # calcPDE.cpp:349-352
    # static_cast<int>(a) == 1 && static_cast<int>(b == 1)   — calcPDE.cpp:383, 391, 395
    AST_LOGICAL_AND: lambda a, b: np.where((_as_int(a) == 1) & (b == 1.0), 1.0, 0.0),
    AST_LOGICAL_OR: lambda a, b: np.where((_as_int(a) == 1) | (b == 1.0), 1.0, 0.0),
    AST_LOGICAL_XOR: lambda a, b: np.where(((_as_int(a) == 1) & (b == 0.0)) | ((_as_int(a) == 0) & (b == 1.0)), 1.0, 0.0),
    AST_RELATIONAL_EQ: lambda a, b: np.where(a == b, 1.0, 0.0), AST_RELATIONAL_GEQ: lambda a, b: np.where(a >= b, 1.0, 0.0),
    AST_RELATIONAL_GT: lambda a, b: np.where(a > b, 1.0, 0.0), AST_RELATIONAL_LEQ: lambda a, b: np.where(a <= b, 1.0, 0.0),
    AST_RELATIONAL_LT: lambda a, b: np.where(a < b, 1.0, 0.0), AST_RELATIONAL_NEQ: lambda a, b: np.where(a != b, 1.0, 0.0),
}


def parse_stream(text):
    """'V name hasDelta' / 'C name value' / 'C # literal' / 'O astType' lines → token tuples."""
    toks = []
    for line in text.splitlines():
        p = line.split()
        if not p:
            continue
        if p[0] == "V":
            toks.append(("V", p[1], int(p[2])))
        elif p[0] == "C":
            toks.append(("C", p[1], float(p[2])))
        else:
            toks.append(("O", int(p[1])))
    return toks


def eval_stream(toks, operand, consts, shape):
    """calcPDE.cpp:244-431 on whole arrays: the same stack discipline for every cell, so the stack is a list of arrays.
    operand(name, has_delta) → array; consts: name → current value (literal tokens carry their value)."""
    # rpStack with its index: a slot keeps what was last stored in it, so a stream that runs the stack empty (every stream with a
    # listed-but-empty operator does) yields the stale bottom slot, exactly like rpStack[st_index] with st_index == 0
    slots = [np.zeros(shape) for _ in range(64)]
    st = 0
    for t in toks:
        if t[0] == "V":
            slots[st] = operand(t[1], t[2])
            st += 1
        elif t[0] == "C":
            v = consts.get(t[1], t[2]) if t[1] != "#" else t[2]
            slots[st] = np.full(shape, v, dtype=np.float64)
            st += 1
        else:
            op = t[1]
            if st > 0:  # if (st_index > 0) st_index--
                st -= 1
            if op in BINARY:
                if st > 0:  # if (st_index > 0) s[st-1] op= s[st]
                    slots[st - 1] = BINARY[op](slots[st - 1], slots[st])
            elif op in UNARY:
                slots[st] = UNARY[op](slots[st])
                st += 1
            # every other operator only pops (calcPDE.cpp:297-302, 379-380, 399-400 ...)
    if st > 0:
        st -= 1
    return slots[st]  # if (st_index > 0) st_index--; rate = s[st_index]


class Species(object):
    def __init__(self, name, u, domain, btype, D, bc=None, adv=None, full=None, dom_d=None, btype_d=None):
        self.name, self.u = name, np.array(u, dtype=np.float64)
        # advected species keep the reference's whole doubled grid (cell values at the even points, CIP face values at
        # the odd ones); u is then a view of its even points
        self.adv = adv          # per axis: uniform velocity or None (adCInfo[a] == 0); None altogether: not advected
        self.full = None
        if adv is not None and any(a is not None for a in adv):
            self.full = np.array(full, dtype=np.float64)
            self.u = self.full[::2, ::2, ::2]
            self.dom_d = np.asarray(dom_d).reshape(-1) == 1          # isDomain == 1 on the doubled grid
            self.btype_d = np.asarray(btype_d, dtype=np.uint8).reshape(-1)
        self.domain = np.asarray(domain, dtype=bool)       # isDomain == 1 of the species' compartment
        self.btype = np.asarray(btype, dtype=np.uint8)     # bits 1..6 = isBofXp, Xm, Yp, Ym, Zp, Zm
        self.D = D                                         # per axis: float or None (diffCInfo[a] == 0)
        self.k = [np.zeros_like(self.u) for _ in range(4)]
        # boundaryInfo: None, or {"Xmax": (kind, value), ...} with kind "neumann" / "dirichlet" / other (ignored)
        self.bc = bc


class MemSpecies(object):
    """A species ON a membrane: like the reference it lives on the whole doubled grid (flat arrays); only the membrane's
    domainIndex points are integrated, the pseudo-membrane points are filled from their neighbours after every step."""

    def __init__(self, name, val, membrane, D, comp=None):
        self.name, self.comp = name, comp
        self.val = np.array(val, dtype=np.float64).reshape(-1)
        self.k = [np.zeros_like(self.val) for _ in range(4)]
        self.membrane = membrane   # Membrane (with the Voronoi data)
        self.D = D                 # diffCInfo[0]: float, array on the doubled grid, or None


class Reaction(object):
    def __init__(self, toks, refs, n_reactants, domain, is_reaction=True, membrane=None, on_membrane=None):
        self.toks, self.refs, self.n_reactants, self.domain, self.is_reaction = toks, refs, n_reactants, domain, is_reaction
        # membrane transport: the Membranes calcMemTransport is called with, in call order; else None
        self.membrane = membrane
        # an ordinary reaction whose first species reference lives on a membrane: that Membrane (reversePolishRK over its points)
        self.on_membrane = on_membrane


class Membrane(object):
    """What calcMemTransport reads of a membrane compartment, all on the reference's doubled grid (flat indices)."""

    def __init__(self, points, is_domain, btype, normals, vol_domains, shape_d, pseudo=(), vor=None):
        self.points = points            # domainIndex of the membrane, in the reference's order
        self.is_domain = is_domain      # isDomain of the membrane (1 membrane point, 2 pseudo membrane)
        self.btype = btype              # per point: bits 1..6 = isBofXp, Xm, Yp, Ym, Zp, Zm
        self.normals = normals          # flat index -> (nx, ny, nz)
        self.vol_domains = vol_domains  # species name -> isDomain of its compartment (flat, doubled grid), membrane species too
        self.shape_d = shape_d          # (Zindex, Yindex, Xindex)
        self.pseudo = list(pseudo)      # pseudoMemIndex
        # voronoiInfo of the points, rows in the order of `points`: (adjacent index, s_ij, d_ij), each [n, 6] with the columns
        # XY[0], YZ[0], XZ[0], XY[1], YZ[1], XZ[1]
        self.vor = vor
        self.points_a = np.array(points, dtype=np.int64)
        self.pseudo_pairs = None        # (pseudo point, neighbour a, neighbour b), built on first use


class RDOracle(object):
    """State + oneStep of the restated path."""

    def __init__(self, species, reactions, consts, delta, dimension, rate_rules=(), fields=None, rules=(), mem_species=()):
        self.sp = {s.name: s for s in species}
        self.order = [s.name for s in species]
        # species on membranes, in the model's species order (they share no array with the volume species, so the two lists can
        # be walked one after the other where the reference walks listOfSpecies once)
        self.msp = {s.name: s for s in mem_species}
        self.morder = [s.name for s in mem_species]
        self.rx = reactions
        # (index of the rule in the model's rule list, domain mask of its variable's compartment): the reference evaluates
        # rInfoList[that index] for a rate rule (spatialsimulator.cpp:1455-1461), which is the rule itself only in a model
        # without reactions
        self.rate_rules = list(rate_rules)
        # variables that are arrays but not integrated (coordinates, parameters and species set by assignments): name -> compact
        # array; ordered assignment rules (orderedARule): (target, domain mask or None for "every point", tokens)
        self.fields = dict(fields or {})
        self.rules = list(rules)
        self.consts = dict(consts)
        self.delta = delta
        self.dimension = dimension
        self.shape = species[0].u.shape

    def _w(self, s, m, dt):
        rk = (0.0, 0.5, 0.5, 1.0)
        if m == 0:
            return s.u
        return s.u + (rk[m] * dt) * s.k[m - 1]  # val + rk[m]*dt*d[(m-1)N + i], calcPDE.cpp:247-248

    def one_step(self, t, dt):
        self.consts["t"] = t * dt  # sim_time = t * dt, spatialsimulator.cpp:1358
        with np.errstate(all="ignore"):
            for m in range(4):
                w = {n: self._w(self.sp[n], m, dt) for n in self.order}

                def operand(name, has_delta):
                    if name in w:
                        return w[name] if has_delta else self.sp[name].u
                    if name in self.fields:
                        return self.fields[name]  # no delta array: the raw value (calcPDE.cpp:249-251)
                    raise NotImplementedError("field operand %s" % name)

                for r in self.rx:  # document order, spatialsimulator.cpp:1374-1390
                    if r is None or not r.is_reaction:
                        continue  # entries without a Reaction (rate rules) are skipped here (:1376-1377)
                    if r.membrane is not None:
                        for mem in r.membrane:  # one calcMemTransport per membrane, spatialsimulator.cpp:1428-1450
                            self._mem_transport(r, mem, m, dt)
                        continue
                    if r.on_membrane is not None:
                        self._mem_reaction(r, m, dt)
                        continue
                    rate = eval_stream(r.toks, operand, self.consts, self.shape)
                    for k, (name, stoich, is_var) in enumerate(r.refs):
                        if not is_var or name not in self.sp:
                            continue
                        km = self.sp[name].k[m]
                        if k < r.n_reactants:
                            km[r.domain] -= (stoich * rate)[r.domain]      # calcPDE.cpp:432-438
                        else:
                            km[r.domain] += (stoich * rate)[r.domain]      # calcPDE.cpp:439-446
                for index, domain in self.rate_rules:  # spatialsimulator.cpp:1455-1461
                    r = self.rx[index]
                    if r is None or r.membrane is not None:
                        raise NotImplementedError("rate rule whose rInfoList slot is a membrane transport")
                    rate = eval_stream(r.toks, operand, self.consts, self.shape)
                    if r.refs and r.refs[0][2] and r.refs[0][0] in self.sp:
                        km = self.sp[r.refs[0][0]].k[m]
                        km[domain] += (r.refs[0][1] * rate)[domain]        # calcPDE.cpp:447-449
                for n in self.order:  # species order, spatialsimulator.cpp:1464-1478
                    s = self.sp[n]
                    self._diffusion(s, w[n], m)
                    self._boundary(s, m)
                for n in self.morder:
                    self._mem_diffusion(self.msp[n], m, dt)
            for n in self.order:  # advection after the RK stages, before the update (spatialsimulator.cpp:1482-1491)
                if self.sp[n].full is not None:
                    self._cip(self.sp[n], dt)
            for n in self.order:  # updateSpecies, spatialsimulator.cpp:1535-1542
                s = self.sp[n]
                inc = dt * (s.k[0] + 2.0 * s.k[1] + 2.0 * s.k[2] + s.k[3]) / 6.0
                s.u[s.domain] += inc[s.domain]
                for m in range(4):
                    s.k[m][s.domain] = 0.0
                # calcBoundary(m = 0) after the k's were zeroed: a non-zero flux is parked in k0 and carried into
                # the next step, whose stage 0 adds it a second time (spatialsimulator.cpp:1540-1542)
                self._boundary(s, 0)
            for n in self.morder:  # updateSpecies over the membrane's domainIndex (no boundary conditions on a membrane)
                s = self.msp[n]
                pts = s.membrane.points_a
                s.val[pts] += (dt * (s.k[0] + 2.0 * s.k[1] + 2.0 * s.k[2] + s.k[3]) / 6.0)[pts]
                for m in range(4):
                    s.k[m][pts] = 0.0
            # updateAssignmentRules (spatialsimulator.cpp:1566-1581): the ordered rules re-evaluated on the new state, a species'
            # rule on the cells of its compartment, a parameter's on every point (reversePolishInitial, calcPDE.cpp:21-222)
            for target, domain, toks in self.rules:
                def now(name, has_delta):
                    if name in self.sp:
                        return self.sp[name].u
                    if name in self.fields:
                        return self.fields[name]
                    raise NotImplementedError("rule operand %s" % name)
                value = eval_stream(toks, now, self.consts, self.shape)
                if target in self.sp:  # a species keeps its arrays (its k's stay zero without reactions); the rule overwrites u
                    self.sp[target].u[domain] = value[domain]
                elif domain is None:
                    self.fields[target] = value.copy()
                else:
                    self.fields[target] = self.fields[target].copy()
                    self.fields[target][domain] = value[domain]
            for n in self.morder:  # pseudo-membrane points, at the end of updateAssignmentRules (spatialsimulator.cpp:1609-1663)
                self._pseudo_fill(self.msp[n])
        return t + dt

    def _mem_reaction(self, r, m, dt):
        """reversePolishRK (calcPDE.cpp:224-451) over the domainIndex of a membrane: every operand is read AT the membrane point
        (val + rk[m]*dt*delta[m-1] when it has a delta array), every variable species reference gains / loses stoich * rate there."""
        rk = (0.0, 0.5, 0.5, 1.0)
        pts = r.on_membrane.points_a

        def operand(name, has_delta):
            if name not in self.msp:
                raise NotImplementedError("operand %s of a reaction on a membrane is not a membrane species" % name)
            s = self.msp[name]
            if m == 0 or not has_delta:
                return s.val[pts]
            return s.val[pts] + (rk[m] * dt) * s.k[m - 1][pts]

        rate = eval_stream(r.toks, operand, self.consts, pts.shape)
        for k, (name, stoich, is_var) in enumerate(r.refs):
            if not is_var:
                continue
            if name not in self.msp:
                raise NotImplementedError("volume species %s in a reaction on a membrane" % name)
            km = self.msp[name].k[m]
            if k < r.n_reactants:
                km[pts] -= stoich * rate
            else:
                km[pts] += stoich * rate

    def _mem_diffusion(self, s, m, dt):
        """calcMemDiffusion, calcPDE.cpp:1484-1581: finite volumes over the Voronoi neighbours of each membrane point.  Per point
        area = sum(d_ij * s_ij) / 2 (2-D) or / 4 (3-D) over the six neighbour slots; per plane that the point's boundary type
        selects, both neighbours j add D * (v_j - v_i) * s_ij / d_ij / area — associated as ((D * dv * s) / d) / area in the
        first stage (:1533) and as (D * ((dv * s) / d)) / area in the later ones (:1555-1557)."""
        if s.D is None:
            return
        rk = (0.0, 0.5, 0.5, 1.0)
        mem = s.membrane
        pts = mem.points_a
        adj, si, di = mem.vor
        area = np.zeros(len(pts))
        for c in range(6):  # j = 0: XY, YZ, XZ; j = 1: XY, YZ, XZ (:1521-1525)
            area = area + di[:, c] * si[:, c]
        area = area / (2.0 if self.dimension == 2 else 4.0)
        D = s.D if np.isscalar(s.D) else np.asarray(s.D).reshape(-1)[pts]  # dcIndex = index for a field-valued coefficient
        bt = mem.btype[pts]
        xp, xm, yp, ym, zp, zm = [(bt >> b) & 1 == 1 for b in range(1, 7)]
        planes = [(0, (xp & xm) | (yp & ym))]
        if self.dimension == 3:
            planes += [(1, (yp & ym) | (zp & zm)), (2, (xp & xm) | (zp & zm))]

        def stage(idx):
            if m == 0:
                return s.val[idx]
            return s.val[idx] + (rk[m] * dt) * s.k[m - 1][idx]

        here = stage(pts)
        km = s.k[m]
        for col, use in planes:
            if not use.any():
                continue
            for j in range(2):
                c = col + 3 * j
                a = adj[use, c]
                if (a < 0).any():
                    raise NotImplementedError("membrane point without a Voronoi neighbour in a plane its boundary type selects")
                dv = stage(a) - here[use]
                Du = D if np.isscalar(D) else D[use]
                if m == 0:
                    km[pts[use]] += ((Du * dv * si[use, c]) / di[use, c]) / area[use]
                else:
                    km[pts[use]] += Du * ((dv * si[use, c]) / di[use, c]) / area[use]

    def _pseudo_fill(self, s):
        """spatialsimulator.cpp:1609-1663: a pseudo-membrane point takes the smaller value of the first pair of its axis
        neighbours that are both membrane points (isDomain == 1), pairs tried in the reference's order."""
        mem = s.membrane
        if mem.pseudo_pairs is None:
            nzd, nyd, nxd = mem.shape_d
            n = nxd * nyd * nzd
            d3 = self.dimension == 3
            first, second, targets = [], [], []
            for index in mem.pseudo:
                z, rem = divmod(index, nxd * nyd)
                y, x = divmod(rem, nxd)

                def on(dx, dy, dz):
                    xx, yy, zz = x + dx, y + dy, z + dz
                    if not (0 <= xx < nxd and 0 <= yy < nyd and 0 <= zz < nzd):
                        return -1
                    f = zz * nyd * nxd + yy * nxd + xx
                    return f if mem.is_domain[f] == 1 else -1
                Xp, Xm, Yp, Ym = on(1, 0, 0), on(-1, 0, 0), on(0, 1, 0), on(0, -1, 0)
                Zp, Zm = (on(0, 0, 1), on(0, 0, -1)) if d3 else (-1, -1)
                order = [(Xp, Xm), (Yp, Ym), (Zp, Zm), (Xp, Yp), (Xp, Ym), (Xp, Zp), (Xp, Zm), (Xm, Yp), (Xm, Ym), (Xm, Zp), (Xm, Zm),
                         (Yp, Zp), (Yp, Zm), (Ym, Zp), (Ym, Zm)]
                for a, b in order:
                    if a >= 0 and b >= 0:
                        targets.append(index)
                        first.append(a)
                        second.append(b)
                        break
            del n
            mem.pseudo_pairs = (np.array(targets, dtype=np.int64), np.array(first, dtype=np.int64), np.array(second, dtype=np.int64))
        t, a, b = mem.pseudo_pairs
        if len(t):
            s.val[t] = np.minimum(s.val[a], s.val[b])

    def _mem_transport(self, r, mem, m, dt):
        """calcMemTransport, calcPDE.cpp:1036-1482, for volume species on either side of a membrane.  Per membrane point
        (not on the frame, isDomain == 1): the rate law is evaluated with every volume operand taken from the adjacent cell
        of whichever side lies in that species' domain, extrapolated 1.5*v(+-1) - 0.5*v(+-3) when the next cell is in the
        domain too (x, then y, then z: later axes overwrite earlier ones, :1107-1213); then reactants lose and products gain
        fabs(n_a) * stoich * rate / delta_a in their cells next to the point, along the axis of the point's boundary type
        (:1394-1479).  Kept as in the reference: the stack outlives the points, and after the evaluation
        `if (st_index > 1) st_index--` (:1392-1393) makes the rate rpStack[1] for a well-formed law (the stale right operand
        of the last binary operator), not the result in rpStack[0]."""
        import math
        nzd, nyd, nxd = mem.shape_d
        n_pts = nxd * nyd * nzd
        rk = (0.0, 0.5, 0.5, 1.0)
        stack = [0.0] * 64

        def cell(flat):  # even point -> compact index
            z, rem = divmod(flat, nxd * nyd)
            y, x = divmod(rem, nxd)
            return (z // 2, y // 2, x // 2)

        def stage_value(sp, flat):  # variable (+ rk[m]*dt*delta[m-1]) at an even point
            c = cell(flat)
            if m == 0:
                return float(sp.u[c])
            return float(sp.u[c]) + rk[m] * dt * float(sp.k[m - 1][c])

        for index in mem.points:
            z, rem = divmod(index, nxd * nyd)
            y, x = divmod(rem, nxd)
            on_frame = (x * y * z == 0 or x == nxd - 1 or y == nyd - 1 or z == nzd - 1) if self.dimension == 3 else \
                       (x * y == 0 or x == nxd - 1 or y == nyd - 1)
            if on_frame or mem.is_domain[index] != 1:
                continue
            strides = (1, nxd, nxd * nyd)
            st = 0
            for t in r.toks:
                if t[0] == "V":
                    name, has_delta = t[1], t[2]
                    if name in self.msp and has_delta:  # a symbol on the membrane: its value at the point (:1214-1218)
                        ms = self.msp[name]
                        stack[st] = float(ms.val[index]) if m == 0 else float(ms.val[index]) + rk[m] * dt * float(ms.k[m - 1][index])
                        st += 1
                        continue
                    if name not in self.sp or not has_delta:
                        raise NotImplementedError("membrane-transport operand %s" % name)
                    sp = self.sp[name]
                    dom = mem.vol_domains[name]
                    for a in range(self.dimension):  # x, y (, z): a later axis overwrites an earlier one
                        p1, p3, m1, m3 = index + strides[a], index + 3 * strides[a], index - strides[a], index - 3 * strides[a]
                        if dom[p1] == 1:
                            if p3 < n_pts and dom[p3] == 1:
                                stack[st] = 1.5 * stage_value(sp, p1) - 0.5 * stage_value(sp, p3)
                            else:
                                stack[st] = stage_value(sp, p1)
                        elif dom[m1] == 1:
                            if m3 < n_pts and dom[m3] == 1:
                                stack[st] = 1.5 * stage_value(sp, m1) - 0.5 * stage_value(sp, m3)
                            else:
                                stack[st] = stage_value(sp, m1)
                    st += 1
                elif t[0] == "C":
                    stack[st] = self.consts.get(t[1], t[2]) if t[1] != "#" else t[2]
                    st += 1
                else:
                    op = t[1]
                    st -= 1
                    a_, b_ = stack[st - 1], stack[st]
                    if op == AST_PLUS:
                        stack[st - 1] = a_ + b_
                    elif op == AST_MINUS:
                        stack[st - 1] = a_ - b_
                    elif op == AST_TIMES:
                        stack[st - 1] = a_ * b_
                    elif op == AST_DIVIDE:
                        stack[st - 1] = a_ / b_ if b_ != 0.0 else math.copysign(math.inf, a_) * math.copysign(1.0, b_) if a_ != 0.0 else math.nan
                    elif op in (AST_POWER, AST_FUNCTION_POWER):
                        stack[st - 1] = math.pow(a_, b_)
                    elif op in BINARY:
                        stack[st - 1] = float(BINARY[op](np.float64(a_), np.float64(b_)))
                    elif op in UNARY:
                        stack[st] = float(UNARY[op](np.float64(b_)))
                        st += 1
                    # every other operator only pops
            if st > 1:
                st -= 1
            rate = stack[st]
            bt = int(mem.btype[index])
            if bt & 0x06:
                a = 0
            elif bt & 0x18:
                a = 1
            elif self.dimension == 3 and bt & 0x60:
                a = 2
            else:
                continue
            n_a = abs(mem.normals[index][a])
            for j, (name, stoich, is_var) in enumerate(r.refs):
                if not is_var or (name not in self.sp and name not in self.msp):
                    continue
                dom = mem.vol_domains[name]
                sign = -1.0 if j < r.n_reactants else 1.0
                amount = n_a * stoich * rate / self.delta[a]
                if name in self.msp:
                    km = self.msp[name].k[m]
                    for nb in (index + strides[a], index - strides[a]):
                        if dom[nb] == 1:
                            km[nb] += sign * amount
                    if dom[index] == 1:  # binding: the species lives on the point itself (:1406-1408)
                        km[index] += sign * (n_a * stoich * rate / (self.delta[a] / 2.0))
                    continue
                sp = self.sp[name]
                if dom[index] == 1:
                    raise NotImplementedError("volume species whose domain holds a membrane point")
                for nb in (index + strides[a], index - strides[a]):  # plus side first, then minus side
                    if dom[nb] == 1:
                        sp.k[m][cell(nb)] += sign * amount

    def _cip(self, s, dt):
        """cipCSLR, calcPDE.cpp:564-869: CIP-CSLR rational-function advection on the doubled grid, one dimension-split
        pass per axis; every pass reads the state before it (val) and writes val_delta, which is then added to the whole
        grid together with the cross-axis face update (val_delta[cell + 2] + val_delta[cell]) / 2.  Uniform and field-valued velocities."""
        import math
        full = s.full
        nzd, nyd, nxd = full.shape
        val = full.reshape(-1)
        n_pts = val.size
        strides, limits = (1, nxd, nxd * nyd), (nxd, nyd, nzd)
        idx = np.flatnonzero(s.dom_d)  # domainIndex: the even points of the species' domain
        coords = (idx % nxd, (idx // nxd) % nyd, idx // (nxd * nyd))
        eps = np.finfo(np.float64).eps  # DBL_EPSILON
        bt = s.btype_d[idx]
        # pow(x, 2): gcc -O2 (the oracle/_ref build, like the reference's own) expands it to x * x

        def in_domain(a, off):
            c = coords[a] + off
            ok = (c >= 0) & (c < limits[a])
            r = np.zeros(idx.shape, dtype=bool)
            r[ok] = s.dom_d[idx[ok] + off * strides[a]]
            return r

        for a in range(self.dimension):
            vd = np.zeros(n_pts)
            u = s.adv[a]
            if u is not None:
                st, dl = strides[a], self.delta[a]
                d4p, d2p, d4m, d2m = in_domain(a, 4), in_domain(a, 2), in_domain(a, -4), in_domain(a, -2)
                p1, p2, p3 = idx + st, idx + 2 * st, idx + 3 * st
                m1, m2, m3 = idx - st, idx - 2 * st, idx - 3 * st
                p3 = np.where(d4p, p3, np.where(d2p, p2, idx))  # :596-603 (uses the uncollapsed +2 index)
                m3 = np.where(d4m, m3, np.where(d2m, m2, idx))
                p1, p2 = np.where(d2p, p1, idx), np.where(d2p, p2, idx)  # :604-611
                m1, m2 = np.where(d2m, m1, idx), np.where(d2m, m2, idx)
                v0, vp1, vp2, vp3 = val[idx], val[p1], val[p2], val[p3]
                vm1, vm2, vm3 = val[m1], val[m2], val[m3]
                # :612-613 a uniform velocity, or the mean of the cell's and its (collapsed) plus neighbour's
                field = not np.isscalar(u)
                ux = (u[idx] + u[p2]) / 2.0 if field else np.full(idx.shape, float(u))
                D = -np.copysign(1.0, ux) * dl
                xi = ux * dt
                pw_mxi, pw_xi = (-xi) * (-xi), xi * xi
                pos = ux >= 0
                g_in, g_out = np.zeros(idx.shape), np.zeros(idx.shape)
                free_p = ((bt >> (1 + 2 * a)) & 1) == 0  # !isBof{X,Y,Z}p
                free_m = ((bt >> (2 + 2 * a)) & 1) == 0
                with np.errstate(all="ignore"):  # (both branches are evaluated everywhere, one is selected)
                    # :614-638 the plus face: flux out for ux >= 0, flux in for ux < 0
                    beta_p = ((np.abs(vp1 - v0) + eps) / (np.abs(v0 - vm1) + eps) - 1.0) / D
                    b_p = ((1.0 + beta_p * D) * v0 - vp1) / D
                    beta_n = ((np.abs(vp1 - vp2) + eps) / (np.abs(vp2 - vp3) + eps) - 1.0) / D
                    b_n = ((1.0 + beta_n * D) * vp2 - vp1) / D
                    beta, b = np.where(pos, beta_p, beta_n), np.where(pos, b_p, b_n)
                    den = 1.0 + beta * (-xi)
                    face = (vp1 + 2 * b * (-xi) + beta * b * pw_mxi) / (den * den) - vp1
                    if field:  # :623-625 correction due to velocity divergence
                        face = face + ((dt / (2 * dl)) * (vp1 + face)) * (u[p3] - u[m1])
                    flux = (-vp1 * xi + b * pw_xi) / (-1.0 + beta * xi)
                    # :639-648 the minus face: flux in for ux >= 0, flux out for ux < 0
                    beta2_p = ((np.abs(vm1 - vm2) + eps) / (np.abs(vm2 - vm3) + eps) - 1.0) / D
                    b2_p = ((1.0 + beta2_p * D) * vm2 - vm1) / D
                    beta2_n = ((np.abs(vm1 - v0) + eps) / (np.abs(v0 - vp1) + eps) - 1.0) / D
                    b2_n = ((1.0 + beta2_n * D) * v0 - vm1) / D
                    beta2, b2 = np.where(pos, beta2_p, beta2_n), np.where(pos, b2_p, b2_n)
                    flux2 = (-vm1 * xi + b2 * pw_xi) / (-1.0 + beta2 * xi)
                vd[p1[free_p]] = face[free_p]
                g_out[free_p & pos] = flux[free_p & pos]
                g_in[free_p & ~pos] = -flux[free_p & ~pos]
                g_in[free_m & pos] = flux2[free_m & pos]
                g_out[free_m & ~pos] = -flux2[free_m & ~pos]
                vd[idx] = (g_in - g_out) / dl  # :650 (after the face value: a collapsed face index is the cell itself)
            # :653-673: faces of the other axes move with the mean of the two cells they separate, then val += val_delta
            for b_ax in range(self.dimension):
                if b_ax == a:
                    continue
                ok = ((bt >> (1 + 2 * b_ax)) & 1) == 0
                c = idx[ok]
                val[c + strides[b_ax]] += (vd[c + 2 * strides[b_ax]] + vd[c]) / 2.0
            val += vd

    def _boundary(self, s, m):
        """calcBoundary, calcPDE.cpp:871-1034, on compact coordinates.  The reference visits every even point c and
        tests its +/-1 neighbours p (target) and +/-2 neighbours pp: p in the domain, pp not (or off the array / on a
        wrapped odd row, which no volume domain contains) -> Neumann: delta[m][p] += -2.0*(-bc)/delta_a (:994), Dirichlet:
        value[p] = bc (:995).  Quirks kept: the Xm test is gated by desc.Y.mOutside (:997: skipped for c on row Y = 0 of
        plane Z = 0); coordinateDesc::set stores Z.p = xp and Z.m = xm (mystruct.h:186-187), so the Z tests hit the X
        neighbour of c and look at c's Z+-2 point; Dirichlet-Zm writes value[Z.mm] (:1028).  A target receives at most one
        contribution per side; their order follows the visiting order of the c's (Z, Y, X ascending; Xp, Xm, Yp, Ym, Zp,
        Zm at each c): Ymax (c = p - row), Xmax and Zmax (c = p - 1), Xmin and Zmin (c = p + 1), Ymin (c = p + row)."""
        if not s.bc:
            return
        dom = s.domain
        nzc, nyc, nxc = dom.shape

        def shifted(di, dj, dk):
            """dom at (i + di, j + dj, k + dk), False off the grid"""
            out = np.zeros_like(dom)
            src = [slice(max(0, d), n + min(0, d)) for d, n in ((dk, nzc), (dj, nyc), (di, nxc))]
            dst = [slice(max(0, -d), n + min(0, -d)) for d, n in ((dk, nzc), (dj, nyc), (di, nxc))]
            out[tuple(dst)] = dom[tuple(src)]
            return out

        K, J, I = np.indices(dom.shape)
        plan = []  # (side, target mask, axis delta)
        if self.dimension >= 2:
            plan.append(("Ymax", dom & (J >= 1) & ~shifted(0, 1, 0), self.delta[1]))
        plan.append(("Xmax", dom & (I >= 1) & ~shifted(1, 0, 0), self.delta[0]))
        if self.dimension >= 3:
            plan.append(("Zmax", dom & (I >= 1) & ~shifted(-1, 0, 2), self.delta[2]))
        plan.append(("Xmin", dom & (I <= nxc - 2) & ~shifted(-1, 0, 0) & ~((K == 0) & (J == 0)), self.delta[0]))
        if self.dimension >= 3:
            plan.append(("Zmin", dom & (I <= nxc - 2) & ~shifted(1, 0, -2), self.delta[2]))
        if self.dimension >= 2:
            plan.append(("Ymin", dom & (J <= nyc - 2) & ~shifted(0, -1, 0), self.delta[1]))
        for side, mask, dl in plan:
            kind, value = s.bc.get(side, (None, 0.0))
            if kind == "neumann":
                s.k[m][mask] += -2.0 * (-value) / dl
            elif kind == "dirichlet":
                if side == "Zmin":  # value[Z.mm]: the cell two planes below c = p + 1; off the array for k < 2
                    kk, jj, ii = np.nonzero(mask)
                    keep = kk - 2 >= 0
                    s.u[kk[keep] - 2, jj[keep], ii[keep] + 1] = value
                else:
                    s.u[mask] = value

    def _diffusion(self, s, w, m):
        """calcPDE.cpp:484-559: Xp, Xm, Yp, Ym, Zp, Zm, each skipped where the boundary-type flag is set."""
        axes = [(2, self.delta[0], 1, 2), (1, self.delta[1], 3, 4), (0, self.delta[2], 5, 6)]  # (array axis, delta, bit+, bit-)
        for a, (ax, dl, bp, bm) in enumerate(axes):
            if s.D[a] is None or (a == 2 and self.dimension <= 2):
                continue
            d2 = np.power(dl, 2)
            for sign, bit in ((+1, bp), (-1, bm)):
                nb = np.roll(w, -sign, axis=ax)  # neighbour value; wrapped entries are masked by the boundary flag
                ok = s.domain & ((s.btype >> bit) & 1 == 0)
                # the grid edge always carries the flag in a domain cell; guard anyway
                idx = [slice(None)] * 3
                idx[ax] = -1 if sign > 0 else 0
                ok = ok.copy()
                ok[tuple(idx)] = False
                term = s.D[a] * (nb - w) / d2
                s.k[m][ok] += term[ok]


def model_parameters(xml_text):
    """Diffusion coefficients (species → [Dx, Dy, Dz] parameter ids), uniform values, compartments."""
    root = ET.fromstring(xml_text)
    values, diff, features, bcs, advs = {}, {}, set(), {}, {}
    for p in root.iter(CORE + "parameter"):
        pid = p.get("id")
        if p.get("value") is not None:
            values[pid] = float(p.get("value"))
        dc = p.find(SPATIAL + "diffusionCoefficient")
        if dc is not None:
            var = dc.get(SPATIAL + "variable") or dc.get("variable")
            ref = dc.get(SPATIAL + "coordinateReference1") or dc.get(SPATIAL + "coordinateIndex")
            kind = dc.get(SPATIAL + "type")
            slots = diff.setdefault(var, [None, None, None])
            if ref in ("cartesianX", "0"):
                slots[0] = pid
            elif ref in ("cartesianY", "1"):
                slots[1] = pid
            elif ref in ("cartesianZ", "2"):
                slots[2] = pid
            else:  # no coordinate reference: every axis (setInfoFunction.cpp:216-227)
                for a in range(3):
                    slots[a] = pid
            del kind
        ac = p.find(SPATIAL + "advectionCoefficient")
        if ac is not None:
            var = ac.get(SPATIAL + "variable")
            ref = ac.get(SPATIAL + "coordinate")
            axis = {"cartesianX": 0, "cartesianY": 1, "cartesianZ": 2}.get(ref)
            if axis is None:
                features.add("advection coefficient without a coordinate")
            else:
                advs.setdefault(var, [None, None, None])[axis] = pid
        bc = p.find(SPATIAL + "boundaryCondition")
        if bc is not None:
            kind = (bc.get(SPATIAL + "type") or "").lower()
            var = bc.get(SPATIAL + "variable")
            side = bc.get(SPATIAL + "coordinateBoundary")
            kind = "dirichlet" if kind in ("dirichlet", "value") else "neumann" if kind in ("neumann", "flux") else kind
            bcs.setdefault(var, {})[side] = (kind, pid)
    for c in root.iter(CORE + "compartment"):
        values[c.get("id")] = float(c.get("size")) if c.get("size") is not None else 1.0
    return values, diff, features, bcs, advs


def from_simulator(sim, xml_text, dims):
    """Build the oracle state from the HOST set-up of a (host-only or device) SpatialSimulator: masks, boundary types,
    token streams and initial fields — each of which the CPU tests pin bit for bit to the reference separately."""
    nx, ny, nz = dims

    def even(a):
        return np.ascontiguousarray(a.reshape(2 * nz - 1, 2 * ny - 1, 2 * nx - 1)[::2, ::2, ::2])

    values, diff, features, bcs, advs = model_parameters(xml_text)
    summary = sim.getSetupText("summary").splitlines()
    head = summary[0].split()
    delta = [float(head[7]), float(head[8]), float(head[9])]
    dimension = int(head[5])
    var_geo, uniform = {}, {}
    for line in summary:
        p = line.split()
        if p[0] == "var":
            if p[3] == "0" and p[7] == "1":  # non-uniform with delta: a spatial species
                var_geo[p[1]] = p[9]
            if p[3] == "1" and "value" in p:
                uniform[p[1]] = float(p[p.index("value") + 1])
    transport_rxn = set()
    mem_adjacent = {}
    for line in summary:
        p = line.split()
        if p[0] == "rxn" and p[4] == "1":
            transport_rxn.add(int(p[1]))
        if p[0] == "geo" and p[5] == "0":  # geo <compartment> domainType <dt> isVol 0 ... adjacent <a> <b>
            mem_adjacent[p[1]] = (p[p.index("adjacent") + 1], p[p.index("adjacent") + 2])
    if features:
        raise NotImplementedError("not restated here: " + ", ".join(sorted(features)))
    shape_d = (2 * nz - 1, 2 * ny - 1, 2 * nx - 1)
    n_d = shape_d[0] * shape_d[1] * shape_d[2]
    # species on membranes are kept apart from the volume species
    mem_geo = {n: c for n, c in var_geo.items() if c in mem_adjacent}
    var_geo = {n: c for n, c in var_geo.items() if c not in mem_adjacent}
    if not var_geo:
        raise NotImplementedError("a model without volume species")
    flat_domain = {}

    def domain_of(comp):  # isDomain of a compartment on the doubled grid (a copy: the simulator hands out borrowed pointers)
        if comp not in flat_domain:
            flat_domain[comp] = np.array(sim.getGeometry(comp)).reshape(-1)[:n_d]
        return flat_domain[comp]

    membranes = {}

    def membrane_of(comp):
        """everything calcMemTransport / calcMemDiffusion read of membrane compartment `comp`"""
        if comp in membranes:
            return membranes[comp]
        bt = sim.getBoundaryType(comp)
        normals = {}
        for line in sim.getSetupText("normals/" + comp).splitlines():
            q = line.split()
            normals[int(q[0])] = (float(q[1]), float(q[2]), float(q[3]))
        points = [int(v) for v in sim.getSetupText("memindex/" + comp).split()]
        pseudo = [int(v) for v in sim.getSetupText("pseudo/" + comp).split()]
        rows = [line.split() for line in sim.getSetupText("voronoi/" + comp).splitlines()]
        vor = None
        if rows:
            if len(rows) != len(points):
                raise NotImplementedError("voronoi rows do not match the membrane's points")
            vor = (np.array([[int(v) for v in q[0:6]] for q in rows], dtype=np.int64),
                   np.array([[float(v) for v in q[6:12]] for q in rows]), np.array([[float(v) for v in q[12:18]] for q in rows]))
        doms = {n: domain_of(c) for n, c in list(var_geo.items()) + list(mem_geo.items())}
        membranes[comp] = Membrane(points, domain_of(comp), sum((bt[:, k].astype(np.uint8) << (k + 1)) for k in range(6)).astype(np.uint8),
                                   normals, doms, shape_d, pseudo, vor)
        return membranes[comp]

    mem_species = []
    for name, comp in mem_geo.items():
        pid = diff.get(name, [None, None, None])[0]  # calcMemDiffusion reads diffCInfo[0] only
        D = None if pid is None else uniform[pid] if pid in uniform else np.array(sim.getVariable(pid))[:n_d]
        if name in bcs or name in advs:
            raise NotImplementedError("boundary condition / advection on membrane species %s" % name)
        mem_species.append(MemSpecies(name, np.array(sim.getVariable(name))[:n_d], membrane_of(comp), D, comp))
        if D is not None and mem_species[-1].membrane.vor is None:
            raise NotImplementedError("membrane %s has no Voronoi data" % comp)
    masks, btypes = {}, {}
    for comp in set(var_geo.values()):
        masks[comp] = even(sim.getGeometry(comp)) == 1
        bt = sim.getBoundaryType(comp)
        btypes[comp] = even(sum((bt[:, k].astype(np.uint8) << (k + 1)) for k in range(6)).astype(np.uint8))
    species = []
    for name, comp in var_geo.items():
        # uniform coefficient: its value; field-valued: the coefficient at the cell itself (calcPDE.cpp:487, dcIndex = index)
        D = []
        for pid in diff.get(name, [None, None, None]):
            if pid is None:
                D.append(None)
            elif pid in uniform:
                D.append(uniform[pid])
            else:
                D.append(even(np.asarray(sim.getVariable(pid))))
        bc = None
        if name in bcs:
            bc = {}
            for side, (kind, pid) in bcs[name].items():
                if pid not in uniform:
                    raise NotImplementedError("field-valued boundary condition %s" % pid)
                bc[side] = (kind, uniform[pid])
        adv = None
        extra = {}
        if name in advs:
            adv = []
            for pid in advs[name]:
                # uniform velocity: its value; field-valued: the parameter on every point of the doubled grid (the face-averaged
                # velocity and the divergence correction read it between the cells too)
                if pid is None:
                    adv.append(None)
                elif pid in uniform:
                    adv.append(uniform[pid])
                else:
                    adv.append(np.array(sim.getVariable(pid))[:(2 * nx - 1) * (2 * ny - 1) * (2 * nz - 1)])
            bt = sim.getBoundaryType(comp)
            extra = {"full": np.asarray(sim.getVariable(name)).reshape(2 * nz - 1, 2 * ny - 1, 2 * nx - 1),
                     "dom_d": sim.getGeometry(comp), "btype_d": sum((bt[:, k].astype(np.uint8) << (k + 1)) for k in range(6))}
        species.append(Species(name, even(sim.getVariable(name)), masks[comp], btypes[comp], D, bc, adv, **extra))
    reactions = []
    i = 0
    while True:
        text = sim.getSetupText("rxn/%d" % i)
        refs_text = sim.getSetupText("refs/%d" % i)
        if not text and not refs_text:
            break
        refs = []
        for line in refs_text.splitlines():
            p = line.split()
            refs.append((p[0], float(p[1]), p[2] == "1"))
        n_reactants = _count_reactants(xml_text, i)
        if n_reactants is None:  # an entry past the reactions: a rate rule (setRateRuleInfo, setInfoFunction.cpp:448-494)
            reactions.append(Reaction(parse_stream(text), refs, 1, None, is_reaction=False))
            i += 1
            continue
        dom = None
        on_mem = None
        for name, _, _ in refs:  # geometry of the first species reference that has one (spatialsimulator.cpp:1381-1390)
            if name in var_geo:
                dom = masks[var_geo[name]]
                break
            if name in mem_geo:
                on_mem = membrane_of(mem_geo[name])
                break
        if i in transport_rxn:
            # calcMemTransport over the membrane between the compartments of the first reactant and the first product when both
            # are volumes (spatialsimulator.cpp:1428-1438); when exactly one of them is a volume, over the membrane the other one
            # lives on (binding / unbinding, :1440-1450)
            names = [name for name, _, _ in refs]
            geo_of = dict(var_geo)
            geo_of.update(mem_geo)
            if n_reactants < 1 or len(names) <= n_reactants or names[0] not in geo_of or names[n_reactants] not in geo_of:
                raise NotImplementedError("membrane transport between species without a geometry")
            rg, pg = geo_of[names[0]], geo_of[names[n_reactants]]
            calls = []
            for c, adj in mem_adjacent.items():  # geoInfoList order
                if adj == (rg, pg) or adj == (pg, rg):
                    calls.append(membrane_of(c))
                    break
            r_vol, p_vol = rg not in mem_adjacent, pg not in mem_adjacent
            if r_vol != p_vol:
                if not r_vol:
                    calls.append(membrane_of(rg))
                if not p_vol:
                    calls.append(membrane_of(pg))
            if not calls and r_vol and p_vol:
                raise NotImplementedError("no membrane between %s" % sorted((rg, pg)))
            reactions.append(Reaction(parse_stream(text), refs, n_reactants, None, membrane=calls))
        elif on_mem is not None:
            reactions.append(Reaction(parse_stream(text), refs, n_reactants, None, on_membrane=on_mem))
        elif dom is not None:
            reactions.append(Reaction(parse_stream(text), refs, n_reactants, dom))
        else:
            reactions.append(None)  # every participant is uniform: keeps the list aligned with rInfoList
        i += 1
    rate_rules = []
    root = ET.fromstring(xml_text)
    rules_el = [r for lr in root.iter(CORE + "listOfRules") for r in lr]
    for index, r in enumerate(rules_el):
        if r.tag != CORE + "rateRule":
            continue
        target = r.get("variable")
        if index >= len(reactions):
            raise NotImplementedError("rate rule %s: the reference indexes its reaction list out of range" % target)
        if target in var_geo:
            rate_rules.append((index, masks[var_geo[target]]))
    # arrays that are not integrated: every non-uniform variable without a delta array (coordinates, parameters and species set by
    # initial assignments / assignment rules), at the volume cells
    fields = {}
    comp_of = {}
    for line in summary:
        p = line.split()
        if p[0] == "var" and p[3] == "0" and p[5] == "1" and p[7] == "0":
            arr = sim.getVariable(p[1])
            if arr is not None:
                fields[p[1]] = even(np.array(arr)[:(2 * nx - 1) * (2 * ny - 1) * (2 * nz - 1)])
                comp_of[p[1]] = p[9]
    rules = []
    i = 0
    while True:
        text = sim.getSetupText("arule/%d" % i)
        if not text:
            break
        target, kind, stream = text.split("\n", 2)
        domain = None
        if kind == "species":  # only the cells of the species' compartment (isAllArea false)
            comp = var_geo.get(target, comp_of.get(target))
            if comp in (None, "-"):
                raise NotImplementedError("assignment rule on species %s without a geometry" % target)
            domain = even(sim.getGeometry(comp)) == 1
        rules.append((target, domain, parse_stream(stream)))
        i += 1
    return RDOracle(species, reactions, uniform, delta, dimension, rate_rules, fields, rules, mem_species)


def _count_reactants(xml_text, index):
    root = ET.fromstring(xml_text)
    rx = [r for r in root.iter(CORE + "reaction") if r.get("fast") != "true"]
    if index >= len(rx):
        return None
    lor = rx[index].find(CORE + "listOfReactants")
    return 0 if lor is None else len(list(lor))
