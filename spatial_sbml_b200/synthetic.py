"""Synthetic SBML-spatial models for the sharded / large-grid benchmarks and tests.

BASELINE.json config 5 ("synthetic turing.xml at 8192x8192 and 512^3"): the 2-species Turing model
of the reference's examples/turing.xml (same kinetic laws with literal constants, D = 4 / 0.4,
Neumann-0 boundaries, blob initial condition) and the Brusselator of examples/The_Brusselator.xml,
written for an arbitrary grid with boundaryMax = n-1 so that deltaX = 1, an all-ones background
volume (ordinal 0) and the reaction domain 1 <= x,y(,z) <= n-2 (ordinal 1).  The membrane
compartment of the example files is left out (no species lives on it).
These files are written from scratch here; nothing is read from the reference at run time.
"""

SPATIAL_NS = "http://www.sbml.org/sbml/level3/version1/spatial/version1"
_MATH = 'xmlns="http://www.w3.org/1998/Math/MathML"'


def _cn(v):
    if isinstance(v, int):
        return '<cn type="integer"> %d </cn>' % v
    return "<cn> %r </cn>" % float(v)


def _ci(n):
    return "<ci> %s </ci>" % n


def _apply(op, *args):
    return "<apply><%s/>%s</apply>" % (op, "".join(args))


def _box(axes, lo, hi):
    """and(x>=lo, x<=hi, y>=lo, ...) as one n-ary <and/>, like the example files."""
    terms = []
    for a, h in zip(axes, hi):
        terms.append(_apply("geq", _ci(a), _cn(lo)))
        terms.append(_apply("leq", _ci(a), _cn(h)))
    return _apply("and", *terms)


def _piecewise(value, cond, otherwise):
    return "<piecewise><piece>%s%s</piece><otherwise>%s</otherwise></piecewise>" % (value, cond, otherwise)


def _param(pid, value, inner=""):
    return '<parameter id="%s" value="%r" constant="true">%s</parameter>' % (pid, float(value), inner)


def _transport_params(species, D, axes, bc_values=None, bc_type="Neumann"):
    out = []
    kinds = {"x": "cartesianX", "y": "cartesianY", "z": "cartesianZ"}
    for a in axes:
        out.append(_param("%s_diff_%s" % (species, a), D,
                          '<spatial:diffusionCoefficient spatial:variable="%s" spatial:type="isotropic" '
                          'spatial:coordinateReference1="%s"/>' % (species, kinds[a])))
    for a in axes:
        for side in ("min", "max"):
            b = "%s%s" % (a.upper(), side)
            v = 0.0 if not bc_values else bc_values.get((species, b), 0.0)
            kind = bc_type
            if isinstance(v, tuple):  # (value, "Dirichlet") overrides the kind of one side
                v, kind = v
            out.append(_param("%s_BC_%s" % (species, b), v,
                              '<spatial:boundaryCondition spatial:variable="%s" spatial:type="%s" '
                              'spatial:coordinateBoundary="%s"/>' % (species, kind, b)))
    return out


def _reaction(rid, reactants, products, math):
    def refs(tag, lst):
        if not lst:
            return ""
        return "<%s>%s</%s>" % (tag, "".join('<speciesReference species="%s" stoichiometry="%r" constant="true"/>' % (s, float(st))
                                               for s, st in lst), tag)
    return ('<reaction id="%s" reversible="false" fast="false">%s%s<kineticLaw><math %s>%s</math></kineticLaw></reaction>'
            % (rid, refs("listOfReactants", reactants), refs("listOfProducts", products), _MATH, math))


def _geometry(axes, dims, lo, hi):
    kinds = {"x": "cartesianX", "y": "cartesianY", "z": "cartesianZ"}
    cc = "".join('<spatial:coordinateComponent spatial:id="%s" spatial:type="%s">'
                 '<spatial:boundaryMin spatial:id="%smin" spatial:value="0"/>'
                 '<spatial:boundaryMax spatial:id="%smax" spatial:value="%d"/></spatial:coordinateComponent>'
                 % (a, kinds[a], a.upper(), a.upper(), n - 1) for a, n in zip(axes, dims))
    inside = _piecewise(_cn(1), _box(axes, lo, hi), _cn(0))
    return ('<spatial:geometry spatial:id="geometry" spatial:coordinateSystem="cartesian">'
            '<spatial:listOfCoordinateComponents>%s</spatial:listOfCoordinateComponents>'
            '<spatial:listOfDomainTypes><spatial:domainType spatial:id="dt_in" spatial:spatialDimensions="3"/>'
            '<spatial:domainType spatial:id="dt_out" spatial:spatialDimensions="3"/></spatial:listOfDomainTypes>'
            '<spatial:listOfDomains><spatial:domain spatial:id="dom_in" spatial:domainType="dt_in"/>'
            '<spatial:domain spatial:id="dom_out" spatial:domainType="dt_out"/></spatial:listOfDomains>'
            '<spatial:listOfGeometryDefinitions><spatial:analyticGeometry spatial:id="ag" spatial:isActive="true">'
            '<spatial:listOfAnalyticVolumes>'
            '<spatial:analyticVolume spatial:id="av_in" spatial:functionType="layered" spatial:ordinal="1" spatial:domainType="dt_in">'
            '<math %s>%s</math></spatial:analyticVolume>'
            '<spatial:analyticVolume spatial:id="av_out" spatial:functionType="layered" spatial:ordinal="0" spatial:domainType="dt_out">'
            '<math %s>%s</math></spatial:analyticVolume>'
            '</spatial:listOfAnalyticVolumes></spatial:analyticGeometry></spatial:listOfGeometryDefinitions>'
            '</spatial:geometry>' % (cc, _MATH, inside, _MATH, _cn(1)))


def _document(model_id, axes, dims, species, params, assignments, reactions, lo=1, hi=None, rules=None):
    hi = hi or [n - 2 for n in dims]
    comps = ('<compartment id="inside" spatialDimensions="3" size="1" constant="true">'
             '<spatial:compartmentMapping spatial:id="m_in" spatial:domainType="dt_in" spatial:unitSize="1"/></compartment>'
             '<compartment id="outside" spatialDimensions="3" size="1" constant="true">'
             '<spatial:compartmentMapping spatial:id="m_out" spatial:domainType="dt_out" spatial:unitSize="1"/></compartment>')
    sp = "".join('<species id="%s" compartment="inside" hasOnlySubstanceUnits="false" boundaryCondition="false" '
                 'constant="false" spatial:isSpatial="true"%s/>' % (s, (' initialConcentration="%r"' % float(c)) if c is not None else "")
                 for s, c in species)
    coords = "".join(_param(a, 0.0, '<spatial:spatialSymbolReference spatial:spatialRef="%s"/>' % a) for a in axes)
    ia = "".join('<initialAssignment symbol="%s"><math %s>%s</math></initialAssignment>' % (s, _MATH, m) for s, m in assignments)
    return ('<?xml version="1.0" encoding="UTF-8"?>\n'
            '<sbml xmlns="http://www.sbml.org/sbml/level3/version1/core" xmlns:spatial="%s" level="3" version="1" '
            'spatial:required="true"><model id="%s"><listOfCompartments>%s</listOfCompartments>'
            '<listOfSpecies>%s</listOfSpecies><listOfParameters>%s%s</listOfParameters>%s%s'
            '%s%s</model></sbml>\n'
            % (SPATIAL_NS, model_id, comps, sp, coords, "".join(params),
               ("<listOfInitialAssignments>%s</listOfInitialAssignments>" % ia) if ia else "",
               ("<listOfRules>%s</listOfRules>" % "".join(rules)) if rules else "",
               ("<listOfReactions>%s</listOfReactions>" % "".join(reactions)) if reactions else "", _geometry(axes, dims, lo, hi)))


def turing(nx, ny, nz=None, blob=True, bc_values=None, D=(4.0, 0.4)):
    """2-species Turing model (examples/turing.xml) on an nx x ny (x nz) grid; dt 0.07 in the reference's annotation."""
    dims = [nx, ny] + ([nz] if nz else [])
    axes = ["x", "y", "z"][:len(dims)]
    params = _transport_params("species_1", D[0], axes, bc_values) + _transport_params("species_2", D[1], axes, bc_values)
    if blob:
        cond = []
        rng = {"x": (0.20, 0.25), "y": (0.70, 0.85), "z": (0.40, 0.60)}
        for a, n in zip(axes, dims):
            L = n - 1
            cond.append(_apply("gt", _ci(a), _cn(rng[a][0] * L)))
            cond.append(_apply("lt", _ci(a), _cn(rng[a][1] * L)))
        ia1 = _piecewise(_cn(20.0), _apply("and", *cond), _cn(0.5))
    else:
        ia1 = _cn(0.5)
    assignments = [("species_1", ia1), ("species_2", _cn(0.1))]
    rx = [
        _reaction("reaction_1", [("species_1", 1)], [("species_2", 1)], _apply("times", _cn(1.0), _ci("species_1"), _ci("species_2"))),
        _reaction("reaction_2", [], [("species_1", 1)], _cn(0.12)),
        _reaction("reaction_3", [], [("species_2", 1)], _cn(0.06)),
        _reaction("reaction_4", [("species_2", 1)], [],
                  _apply("times", _cn(0.5), _ci("species_2"), _apply("divide", _cn(1.0), _apply("plus", _cn(0.1), _ci("species_2"))))),
    ]
    return _document("synthetic_turing", axes, dims, [("species_1", None), ("species_2", None)], params, assignments, rx)


def brusselator(nx, ny, nz=None, bc_values=None, perturb=False):
    """Brusselator (examples/The_Brusselator.xml: k = 0.1, A = 2.5, B = 5.24, D = 0.16 / 0.8); dt 0.1."""
    dims = [nx, ny] + ([nz] if nz else [])
    axes = ["x", "y", "z"][:len(dims)]
    params = _transport_params("X", 0.16, axes, bc_values) + _transport_params("Y", 0.8, axes, bc_values)
    params += [_param("k", 0.1), _param("A", 2.5), _param("B", 5.24)]
    assignments = []
    species = [("X", 2.5), ("Y", 2.0)]
    if perturb:
        # smooth, non-uniform start so that diffusion terms are exercised: X = 2.5 + 0.01*x
        species = [("X", None), ("Y", 2.0)]
        assignments = [("X", _apply("plus", _cn(2.5), _apply("times", _cn(0.01), _ci("x"))))]
    rx = [
        _reaction("J0", [], [("X", 1)], _apply("times", _ci("k"), _ci("A"))),
        _reaction("J1", [("Y", 1)], [("X", 1)], _apply("times", _ci("k"), _ci("X"), _ci("X"), _ci("Y"))),
        _reaction("J2", [("X", 1)], [("Y", 1)], _apply("times", _ci("k"), _ci("X"), _ci("B"))),
        _reaction("J3", [("X", 1)], [], _apply("times", _ci("k"), _ci("X"))),
    ]
    return _document("synthetic_brusselator", axes, dims, species, params, assignments, rx)


def cascade3(nx, ny, nz=None):
    """Three diffusing species in one compartment: the Brusselator pair plus a product Z that X decays into and that is
    degraded with saturating kinetics (a division and a comparison in the law).  Exercises the three- and four-species
    kernel variants, which the reference's own two-species synthetic models leave idle in 3-D.  dt 0.05."""
    dims = [nx, ny] + ([nz] if nz else [])
    axes = ["x", "y", "z"][:len(dims)]
    params = (_transport_params("X", 0.16, axes, None) + _transport_params("Y", 0.8, axes, None) +
              _transport_params("Z", 0.3, axes, None))
    params += [_param("k", 0.1), _param("A", 2.5), _param("B", 5.24), _param("Vz", 0.4), _param("Kz", 0.7)]
    species = [("X", None), ("Y", 2.0), ("Z", None)]
    assignments = [("X", _apply("plus", _cn(2.5), _apply("times", _cn(0.01), _ci("x")))),
                   ("Z", _apply("plus", _cn(0.2), _apply("times", _cn(0.02), _ci("y"))))]
    rx = [
        _reaction("J0", [], [("X", 1)], _apply("times", _ci("k"), _ci("A"))),
        _reaction("J1", [("Y", 1)], [("X", 1)], _apply("times", _ci("k"), _ci("X"), _ci("X"), _ci("Y"))),
        _reaction("J2", [("X", 1)], [("Y", 1)], _apply("times", _ci("k"), _ci("X"), _ci("B"))),
        _reaction("J3", [("X", 1)], [("Z", 2)], _apply("times", _ci("k"), _ci("X"))),
        _reaction("J4", [("Z", 1)], [],
                  _apply("times", _apply("divide", _apply("times", _ci("Vz"), _ci("Z")), _apply("plus", _ci("Kz"), _ci("Z"))),
                         _apply("gt", _ci("Z"), _cn(0.05)))),
    ]
    return _document("synthetic_cascade3", axes, dims, species, params, assignments, rx)


def turing_flux3d(nx, ny, nz):
    """The 3-D Turing model with a non-zero Neumann flux on every side of species_1 and, for species_2, fluxes in x / y
    and Dirichlet values on both z sides: the boundary pass of the reference in 3-D, including its Z-side quirks
    (coordinateDesc stores the X neighbour in the Z slots, Dirichlet-Zmin writes two planes below).  dt 0.02."""
    bc = {("species_1", "Xmin"): 0.3, ("species_1", "Xmax"): -0.1, ("species_1", "Ymin"): 0.2, ("species_1", "Ymax"): 0.05,
          ("species_1", "Zmin"): 0.15, ("species_1", "Zmax"): -0.07,
          ("species_2", "Xmin"): 0.02, ("species_2", "Ymax"): 0.04,
          ("species_2", "Zmax"): (0.12, "Dirichlet"), ("species_2", "Zmin"): (0.09, "Dirichlet")}
    return turing(nx, ny, nz, bc_values=bc)



def turing_bc(nx, ny):
    """2-D Turing model whose boundary pass is live: non-zero Neumann fluxes on three sides of species_1 and a Dirichlet value on
    Ymin of species_2 (the unfused schedule with the flux carried in k0).  dt 0.05."""
    bc = {("species_1", "Xmin"): 0.3, ("species_1", "Xmax"): -0.1, ("species_1", "Ymax"): 0.05,
          ("species_2", "Xmin"): 0.02, ("species_2", "Ymin"): (0.12, "Dirichlet")}
    return turing(nx, ny, bc_values=bc)


def _membrane_geometry(axes, dims, inside_math):
    """outside (ordinal 0, everything) / cell (ordinal 1, `inside_math`) / the membrane between them, declared the way the
    reference's transport.xml declares its compartments (domain types, domains, adjacentDomains)."""
    kinds = {"x": "cartesianX", "y": "cartesianY", "z": "cartesianZ"}
    cc = "".join('<spatial:coordinateComponent spatial:id="%s" spatial:type="%s">'
                 '<spatial:boundaryMin spatial:id="%smin" spatial:value="0"/>'
                 '<spatial:boundaryMax spatial:id="%smax" spatial:value="%d"/></spatial:coordinateComponent>'
                 % (a, kinds[a], a.upper(), a.upper(), n - 1) for a, n in zip(axes, dims))
    return ('<spatial:geometry spatial:id="geometry" spatial:coordinateSystem="cartesian">'
            '<spatial:listOfCoordinateComponents>%s</spatial:listOfCoordinateComponents>'
            '<spatial:listOfDomainTypes><spatial:domainType spatial:id="dt_cell" spatial:spatialDimensions="3"/>'
            '<spatial:domainType spatial:id="dt_out" spatial:spatialDimensions="3"/>'
            '<spatial:domainType spatial:id="dt_mem" spatial:spatialDimensions="2"/></spatial:listOfDomainTypes>'
            '<spatial:listOfDomains><spatial:domain spatial:id="dom_cell" spatial:domainType="dt_cell"/>'
            '<spatial:domain spatial:id="dom_out" spatial:domainType="dt_out"/>'
            '<spatial:domain spatial:id="dom_mem" spatial:domainType="dt_mem"/></spatial:listOfDomains>'
            '<spatial:listOfAdjacentDomains>'
            '<spatial:adjacentDomains spatial:id="ad0" spatial:domain1="dom_mem" spatial:domain2="dom_cell"/>'
            '<spatial:adjacentDomains spatial:id="ad1" spatial:domain1="dom_mem" spatial:domain2="dom_out"/>'
            '</spatial:listOfAdjacentDomains>'
            '<spatial:listOfGeometryDefinitions><spatial:analyticGeometry spatial:id="ag" spatial:isActive="true">'
            '<spatial:listOfAnalyticVolumes>'
            '<spatial:analyticVolume spatial:id="av_cell" spatial:functionType="layered" spatial:ordinal="1" spatial:domainType="dt_cell">'
            '<math %s>%s</math></spatial:analyticVolume>'
            '<spatial:analyticVolume spatial:id="av_out" spatial:functionType="layered" spatial:ordinal="0" spatial:domainType="dt_out">'
            '<math %s>%s</math></spatial:analyticVolume>'
            '</spatial:listOfAnalyticVolumes></spatial:analyticGeometry></spatial:listOfGeometryDefinitions>'
            '</spatial:geometry>' % (cc, _MATH, inside_math, _MATH, _cn(1)))


def membrane(nx, ny, nz=None, shape="disc", transport=True, mixed=False, surface=True):
    """A cell (disc / ball, or a box with shape="box") in a bath with species ON the membrane between them: what the reference
    calls membrane diffusion (calcMemDiffusion over the Voronoi neighbours of setVoronoiInfo), binding of a volume species to
    a membrane species (the binding branches of calcMemTransport), a reaction between membrane species (reversePolishRK over
    the membrane's points), the pseudo-membrane fill of updateAssignmentRules, and — with transport=True — an ordinary
    membrane transport between the two volumes beside them.  No example file of the reference has a membrane species.
      A (cell, diffuses) --bind: kon*A--> M (membrane, surface diffusion, non-uniform start) --conv: kc*M--> N (membrane)
      M --unbind: koff*M--> A;   A (cell) --leak: kl*A--> E (bath), a transport across the membrane.
    mixed=True: ordinary volume reactions on the transported species as well, listed BEFORE the transports (-> A: g, A ->: kd*A)
    and after them (E ->: ke*E), so that the order in which the reference accumulates the terms of a species (rInfoList order,
    spatialsimulator.cpp:1374-1451) matters to the last bit.  surface=False: no species on the membrane (M, N and their reactions
    are left out), only the transport across it.
    dt 0.05."""
    dims = [nx, ny] + ([nz] if nz else [])
    axes = ["x", "y", "z"][:len(dims)]
    c = [(n - 1) / 2.0 + 0.25 * (i + 1) for i, n in enumerate(dims)]  # slightly off the grid's symmetry axes
    r = 0.31 * (min(dims) - 1)
    if shape == "box":
        inside = _piecewise(_cn(1), _box(axes, int(0.25 * (min(dims) - 1)), [int(0.7 * (n - 1)) for n in dims]), _cn(0))
    else:
        sq = [_apply("power", _apply("minus", _ci(a), _cn(ci)), _cn(2)) for a, ci in zip(axes, c)]
        inside = _piecewise(_cn(1), _apply("leq", _apply("plus", *sq), _cn(r * r)), _cn(0))
    comps = ('<compartment id="cell" spatialDimensions="3" size="1" constant="true">'
             '<spatial:compartmentMapping spatial:id="m_cell" spatial:domainType="dt_cell" spatial:unitSize="1"/></compartment>'
             '<compartment id="bath" spatialDimensions="3" size="1" constant="true">'
             '<spatial:compartmentMapping spatial:id="m_out" spatial:domainType="dt_out" spatial:unitSize="1"/></compartment>'
             '<compartment id="mem" spatialDimensions="2" size="1" constant="true">'
             '<spatial:compartmentMapping spatial:id="m_mem" spatial:domainType="dt_mem" spatial:unitSize="1"/></compartment>')
    def sp(sid, comp, conc):
        return ('<species id="%s" compartment="%s" hasOnlySubstanceUnits="false" boundaryCondition="false" constant="false" '
                'spatial:isSpatial="true"%s/>' % (sid, comp, (' initialConcentration="%r"' % float(conc)) if conc is not None else ""))
    species = sp("A", "cell", None) + sp("E", "bath", 0.25) + (sp("M", "mem", None) + sp("N", "mem", 0.0) if surface else "")
    params = _transport_params("A", 0.8, axes) + _transport_params("E", 0.5, axes)
    if surface:
        params.append(_param("M_diff", 0.35, '<spatial:diffusionCoefficient spatial:variable="M" spatial:type="isotropic"/>'))
        params.append(_param("N_diff", 0.2, '<spatial:diffusionCoefficient spatial:variable="N" spatial:type="isotropic"/>'))
    params += [_param("kon", 0.3), _param("koff", 0.05), _param("kc", 0.11), _param("kl", 0.07)]
    coords = "".join(_param(a, 0.0, '<spatial:spatialSymbolReference spatial:spatialRef="%s"/>' % a) for a in axes)
    ia = [("A", _apply("plus", _cn(1.0), _apply("times", _cn(0.02), _ci("x")))),
          ("M", _apply("plus", _cn(0.5), _apply("times", _cn(0.03), _ci("y")), _apply("times", _cn(0.01), _ci("x"))))]
    if not surface:
        ia = ia[:1]
    ia_xml = "".join('<initialAssignment symbol="%s"><math %s>%s</math></initialAssignment>' % (s, _MATH, m) for s, m in ia)
    rx = [] if not surface else [_reaction("bind", [("A", 1)], [("M", 1)], _apply("times", _ci("kon"), _ci("A"))),
          _reaction("conv", [("M", 1)], [("N", 1)], _apply("times", _ci("kc"), _ci("M"))),
          _reaction("unbind", [("M", 1)], [("A", 1)], _apply("times", _ci("koff"), _ci("M")))]
    if transport:
        rx.append(_reaction("leak", [("A", 1)], [("E", 1)], _apply("times", _ci("kl"), _ci("A"))))
    if mixed:
        params += [_param("g", 0.013), _param("kd", 0.021), _param("ke", 0.017)]
        rx = ([_reaction("growA", [], [("A", 1)], _ci("g")), _reaction("decayA", [("A", 1)], [], _apply("times", _ci("kd"), _ci("A")))]
              + rx + [_reaction("decayE", [("E", 1)], [], _apply("times", _ci("ke"), _ci("E")))])
    return ('<?xml version="1.0" encoding="UTF-8"?>\n'
            '<sbml xmlns="http://www.sbml.org/sbml/level3/version1/core" xmlns:spatial="%s" level="3" version="1" '
            'spatial:required="true"><model id="synthetic_membrane"><listOfCompartments>%s</listOfCompartments>'
            '<listOfSpecies>%s</listOfSpecies><listOfParameters>%s%s</listOfParameters>'
            '<listOfInitialAssignments>%s</listOfInitialAssignments>'
            '<listOfReactions>%s</listOfReactions>%s</model></sbml>\n'
            % (SPATIAL_NS, comps, species, coords, "".join(params), ia_xml, "".join(rx), _membrane_geometry(axes, dims, inside)))


def membrane_mixed(nx, ny, nz=None):
    """membrane() with ordinary reactions on the transported species before and after the transports in the reaction list"""
    return membrane(nx, ny, nz, mixed=True)


def transport_mixed(nx, ny, nz=None):
    """A transport across a membrane without membrane species, with ordinary reactions on the transported species listed before
    (A) and after (E) the transport"""
    return membrane(nx, ny, nz, mixed=True, surface=False)


def membrane_box(nx, ny, nz=None):
    """membrane() around a box-shaped cell: straight membrane segments and corners (the corner branches of the pseudo-membrane fill)."""
    return membrane(nx, ny, nz, shape="box")


_TIME = '<csymbol encoding="text" definitionURL="http://www.sbml.org/sbml/symbols/time"> t </csymbol>'


def _assignment_rule(var, math):
    return '<assignmentRule variable="%s"><math %s>%s</math></assignmentRule>' % (var, _MATH, math)


def _rate_rule(var, math):
    return '<rateRule variable="%s"><math %s>%s</math></rateRule>' % (var, _MATH, math)


def _varparam(pid):
    return '<parameter id="%s" constant="false"/>' % pid


def rules(nx, ny, nz=None):
    """Brusselator kinetics steered by ASSIGNMENT RULES that the reference re-evaluates after every step (updateAssignmentRules,
    spatialsimulator.cpp:1566-1581): `act`, a field that depends on time and position (sin, exp, unary minus), `gain`, a field
    that depends on the species (a division, abs, ln, root, power), `W`, a species defined by a rule, and `late`, a rule listed
    before the rule it depends on (the dependency ordering of spatialsimulator.cpp:1108-1153).  dt 0.05."""
    dims = [nx, ny] + ([nz] if nz else [])
    axes = ["x", "y", "z"][:len(dims)]
    params = _transport_params("X", 0.16, axes) + _transport_params("Y", 0.8, axes)
    params += [_param("k", 0.1), _param("A", 2.5), _param("B", 5.24), _varparam("late"), _varparam("act"), _varparam("gain")]
    species = [("X", None), ("Y", 2.0), ("W", None)]
    assignments = [("X", _apply("plus", _cn(2.5), _apply("times", _cn(0.01), _ci("x"))))]
    act = _apply("plus", _cn(1.0), _apply("times", _cn(0.4), _apply("sin", _apply("times", _cn(0.9), _TIME)),
                                          _apply("exp", _apply("minus", _apply("times", _cn(0.02), _ci("y"))))))
    gain = _apply("divide", _apply("plus", _cn(1.0), _apply("abs", _apply("minus", _ci("X"), _ci("Y")))),
                  _apply("plus", _cn(1.5), _apply("ln", _apply("plus", _cn(1.0), _apply("root", _apply("power", _ci("Y"), _cn(2)))))))
    rules_ = [_assignment_rule("late", _apply("times", _cn(0.5), _ci("act"), _ci("gain"))),
              _assignment_rule("act", act), _assignment_rule("gain", gain),
              _assignment_rule("W", _apply("times", _cn(0.5), _apply("plus", _ci("X"), _ci("Y"))))]
    rx = [
        _reaction("J0", [], [("X", 1)], _apply("times", _ci("k"), _ci("A"), _ci("act"))),
        _reaction("J1", [("Y", 1)], [("X", 1)], _apply("times", _ci("k"), _ci("X"), _ci("X"), _ci("Y"))),
        _reaction("J2", [("X", 1)], [("Y", 1)], _apply("times", _ci("k"), _ci("X"), _ci("B"), _ci("gain"))),
        _reaction("J3", [("X", 1)], [], _apply("times", _ci("k"), _ci("X"), _apply("plus", _cn(0.8), _ci("late")))),
    ]
    return _document("synthetic_rules", axes, dims, species, params, assignments, rx, rules=rules_)


def raterules(nx, ny, nz=None):
    """Two diffusing species driven by RATE RULES only (no reactions, so that the reference's rInfoList[ruleIndex] is the rule
    itself, spatialsimulator.cpp:1455-1461), with the rarely used operators in the laws: power, log to a base, xor / eq / neq,
    piecewise, tanh, cos, a listed-but-empty operator that only pops (floor) and the `time` symbol.  dt 0.05."""
    dims = [nx, ny] + ([nz] if nz else [])
    axes = ["x", "y", "z"][:len(dims)]
    params = _transport_params("P", 0.3, axes) + _transport_params("Q", 0.6, axes)
    params += [_param("a", 0.7), _param("b", 0.2)]
    species = [("P", None), ("Q", None)]
    assignments = [("P", _apply("plus", _cn(1.0), _apply("times", _cn(0.02), _ci("x")))),
                   ("Q", _apply("plus", _cn(0.5), _apply("times", _cn(0.01), _ci("y"))))]
    # dP/dt = a*Q^1.5 - b*P*tanh(P) + 0.05*piecewise(1, xor(P > 1.2, Q > 0.6), 0) + 0.01*cos(t)
    dP = _apply("plus",
                _apply("minus", _apply("times", _ci("a"), _apply("power", _ci("Q"), _cn(1.5))),
                       _apply("times", _ci("b"), _ci("P"), _apply("tanh", _ci("P")))),
                _apply("times", _cn(0.05), _piecewise(_cn(1), _apply("xor", _apply("gt", _ci("P"), _cn(1.2)), _apply("gt", _ci("Q"), _cn(0.6))), _cn(0))),
                _apply("times", _cn(0.01), _apply("cos", _TIME)))
    # dQ/dt = 0.1*log_2(1 + P) - 0.3*Q*(Q != 0.5) + 0.02*(P == Q)
    dQ = _apply("plus",
                _apply("minus", _apply("times", _cn(0.1), "<apply><log/><logbase>%s</logbase>%s</apply>" % (_cn(2), _apply("plus", _cn(1.0), _ci("P")))),
                       _apply("times", _cn(0.3), _ci("Q"), _apply("neq", _ci("Q"), _cn(0.5)))),
                _apply("times", _cn(0.02), _apply("eq", _ci("P"), _ci("Q"))))
    # dR/dt = (0.1 + 0.3) + 0.02*floor(P): floor is listed but empty in the reference's interpreter - it pops its operand and
    # pushes nothing, the stream runs the stack empty and the rate is what was last stored at the bottom of the stack
    dR = _apply("plus", _apply("plus", _cn(0.1), _cn(0.3)), _apply("times", _cn(0.02), _apply("floor", _ci("P"))))
    rules_ = [_rate_rule("P", dP), _rate_rule("Q", dQ), _rate_rule("R", dR)]
    species.append(("R", 1.0))
    params += _transport_params("R", 0.2, axes)
    return _document("synthetic_raterules", axes, dims, species, params, assignments, [], rules=rules_)


def sampled(nx, ny):
    """The Brusselator in a cell whose shape comes from an IMAGE: a SampledFieldGeometry (reference spatialsimulator.cpp:483-604;
    the image rows are stored top-down, so sample (x, y) of the grid is row numSamples2 - 1 - y) with two sampled volumes, an
    L-shaped cell (value 1) in a bath (value 0), and the membrane between them.  dt 0.05."""
    img = [[0] * nx for _ in range(ny)]
    for y in range(ny):
        for x in range(nx):
            in_l = (3 <= x <= nx // 3 and 3 <= y <= ny - 4) or (3 <= x <= nx - 5 and 3 <= y <= ny // 3)
            in_hole = (5 <= x <= 7 and 5 <= y <= 7)
            img[ny - 1 - y][x] = 1 if (in_l and not in_hole) else 0
    samples = " ".join(str(v) for row in img for v in row)
    axes, dims = ["x", "y"], [nx, ny]
    kinds = {"x": "cartesianX", "y": "cartesianY"}
    cc = "".join('<spatial:coordinateComponent spatial:id="%s" spatial:type="%s">'
                 '<spatial:boundaryMin spatial:id="%smin" spatial:value="0"/>'
                 '<spatial:boundaryMax spatial:id="%smax" spatial:value="%d"/></spatial:coordinateComponent>'
                 % (a, kinds[a], a.upper(), a.upper(), n - 1) for a, n in zip(axes, dims))
    geometry = ('<spatial:geometry spatial:id="geometry" spatial:coordinateSystem="cartesian">'
                '<spatial:listOfCoordinateComponents>%s</spatial:listOfCoordinateComponents>'
                '<spatial:listOfDomainTypes><spatial:domainType spatial:id="dt_cell" spatial:spatialDimensions="3"/>'
                '<spatial:domainType spatial:id="dt_out" spatial:spatialDimensions="3"/>'
                '<spatial:domainType spatial:id="dt_mem" spatial:spatialDimensions="2"/></spatial:listOfDomainTypes>'
                '<spatial:listOfDomains><spatial:domain spatial:id="dom_cell" spatial:domainType="dt_cell"/>'
                '<spatial:domain spatial:id="dom_out" spatial:domainType="dt_out"/>'
                '<spatial:domain spatial:id="dom_mem" spatial:domainType="dt_mem"/></spatial:listOfDomains>'
                '<spatial:listOfAdjacentDomains>'
                '<spatial:adjacentDomains spatial:id="ad0" spatial:domain1="dom_mem" spatial:domain2="dom_cell"/>'
                '<spatial:adjacentDomains spatial:id="ad1" spatial:domain1="dom_mem" spatial:domain2="dom_out"/>'
                '</spatial:listOfAdjacentDomains>'
                '<spatial:listOfGeometryDefinitions><spatial:sampledFieldGeometry spatial:id="sfg" spatial:isActive="true" '
                'spatial:sampledField="img"><spatial:listOfSampledVolumes>'
                '<spatial:sampledVolume spatial:id="sv_cell" spatial:domainType="dt_cell" spatial:sampledValue="1"/>'
                '<spatial:sampledVolume spatial:id="sv_out" spatial:domainType="dt_out" spatial:sampledValue="0"/>'
                '</spatial:listOfSampledVolumes></spatial:sampledFieldGeometry></spatial:listOfGeometryDefinitions>'
                '<spatial:listOfSampledFields><spatial:sampledField spatial:id="img" spatial:dataType="uint8" '
                'spatial:numSamples1="%d" spatial:numSamples2="%d" spatial:numSamples3="1" spatial:interpolationType="nearestNeighbor" '
                'spatial:compression="uncompressed" spatial:samplesLength="%d">%s</spatial:sampledField></spatial:listOfSampledFields>'
                '</spatial:geometry>' % (cc, nx, ny, nx * ny, samples))
    comps = ('<compartment id="cell" spatialDimensions="3" size="1" constant="true">'
             '<spatial:compartmentMapping spatial:id="m_cell" spatial:domainType="dt_cell" spatial:unitSize="1"/></compartment>'
             '<compartment id="bath" spatialDimensions="3" size="1" constant="true">'
             '<spatial:compartmentMapping spatial:id="m_out" spatial:domainType="dt_out" spatial:unitSize="1"/></compartment>'
             '<compartment id="mem" spatialDimensions="2" size="1" constant="true">'
             '<spatial:compartmentMapping spatial:id="m_mem" spatial:domainType="dt_mem" spatial:unitSize="1"/></compartment>')
    def sp(sid, comp, conc):
        return ('<species id="%s" compartment="%s" hasOnlySubstanceUnits="false" boundaryCondition="false" constant="false" '
                'spatial:isSpatial="true"%s/>' % (sid, comp, (' initialConcentration="%r"' % float(conc)) if conc is not None else ""))
    species = sp("X", "cell", None) + sp("Y", "cell", 2.0) + sp("E", "bath", 0.1)
    params = _transport_params("X", 0.16, axes) + _transport_params("Y", 0.8, axes) + _transport_params("E", 0.4, axes)
    params += [_param("k", 0.1), _param("A", 2.5), _param("B", 5.24), _param("kl", 0.05)]
    coords = "".join(_param(a, 0.0, '<spatial:spatialSymbolReference spatial:spatialRef="%s"/>' % a) for a in axes)
    ia = '<initialAssignment symbol="X"><math %s>%s</math></initialAssignment>' % (
        _MATH, _apply("plus", _cn(2.5), _apply("times", _cn(0.02), _ci("x"))))
    rx = [_reaction("J0", [], [("X", 1)], _apply("times", _ci("k"), _ci("A"))),
          _reaction("J1", [("Y", 1)], [("X", 1)], _apply("times", _ci("k"), _ci("X"), _ci("X"), _ci("Y"))),
          _reaction("J2", [("X", 1)], [("Y", 1)], _apply("times", _ci("k"), _ci("X"), _ci("B"))),
          _reaction("J3", [("X", 1)], [], _apply("times", _ci("k"), _ci("X"))),
          _reaction("leak", [("Y", 1)], [("E", 1)], _apply("times", _ci("kl"), _ci("Y")))]
    return ('<?xml version="1.0" encoding="UTF-8"?>\n'
            '<sbml xmlns="http://www.sbml.org/sbml/level3/version1/core" xmlns:spatial="%s" level="3" version="1" '
            'spatial:required="true"><model id="synthetic_sampled"><listOfCompartments>%s</listOfCompartments>'
            '<listOfSpecies>%s</listOfSpecies><listOfParameters>%s%s</listOfParameters>'
            '<listOfInitialAssignments>%s</listOfInitialAssignments>'
            '<listOfReactions>%s</listOfReactions>%s</model></sbml>\n'
            % (SPATIAL_NS, comps, species, coords, "".join(params), ia, "".join(rx), geometry))


def fieldcoeff(nx, ny, nz=None):
    """Diffusion coefficients that are FIELDS (a parameter with an initial assignment in x / y: the reference reads
    diffCInfo->value[index] per cell, calcPDE.cpp:486-511) on a grid with non-unit spacing would take the general stencil path;
    here the spacing is 1 and only the coefficients vary.  Brusselator kinetics.  dt 0.05."""
    dims = [nx, ny] + ([nz] if nz else [])
    axes = ["x", "y", "z"][:len(dims)]
    kinds = {"x": "cartesianX", "y": "cartesianY", "z": "cartesianZ"}
    params = []
    ia = [("X", _apply("plus", _cn(2.5), _apply("times", _cn(0.03), _ci("x")))),
          ("Y", _apply("plus", _cn(2.0), _apply("times", _cn(0.02), _ci("y"))))]
    for s_, base in (("X", 0.16), ("Y", 0.5)):
        for k_, a in enumerate(axes):
            pid = "%s_diff_%s" % (s_, a)
            params.append('<parameter id="%s" constant="true"><spatial:diffusionCoefficient spatial:variable="%s" '
                          'spatial:type="isotropic" spatial:coordinateReference1="%s"/></parameter>' % (pid, s_, kinds[a]))
            ia.append((pid, _apply("plus", _cn(base), _apply("times", _cn(0.004 * (k_ + 1)), _ci(axes[(k_ + 1) % len(axes)])))))
        for a in axes:
            for side in ("min", "max"):
                b = "%s%s" % (a.upper(), side)
                params.append(_param("%s_BC_%s" % (s_, b), 0.0, '<spatial:boundaryCondition spatial:variable="%s" spatial:type="Neumann" '
                                                               'spatial:coordinateBoundary="%s"/>' % (s_, b)))
    params += [_param("k", 0.1), _param("A", 2.5), _param("B", 5.24)]
    rx = [_reaction("J0", [], [("X", 1)], _apply("times", _ci("k"), _ci("A"))),
          _reaction("J1", [("Y", 1)], [("X", 1)], _apply("times", _ci("k"), _ci("X"), _ci("X"), _ci("Y"))),
          _reaction("J2", [("X", 1)], [("Y", 1)], _apply("times", _ci("k"), _ci("X"), _ci("B"))),
          _reaction("J3", [("X", 1)], [], _apply("times", _ci("k"), _ci("X")))]
    return _document("synthetic_fieldcoeff", axes, dims, [("X", None), ("Y", None)], params, ia, rx)


def chain6(nx, ny, nz=None):
    """Six diffusing species in one compartment, a linear conversion chain S0 -> S1 -> ... -> S5 -> (degraded) with a feed into S0
    and one nonlinear step: more staged species than the rd_stage kernel variants hold (4), so the first-generation stage kernel's
    8-species instantiation runs.  dt 0.05."""
    dims = [nx, ny] + ([nz] if nz else [])
    axes = ["x", "y", "z"][:len(dims)]
    names = ["S%d" % i for i in range(6)]
    params = []
    for i, n in enumerate(names):
        params += _transport_params(n, 0.1 + 0.07 * i, axes)
    params += [_param("kf", 0.2), _param("kc", 0.15), _param("kd", 0.05)]
    species = [(n, None if i == 0 else 0.1 * (i + 1)) for i, n in enumerate(names)]
    assignments = [("S0", _apply("plus", _cn(1.0), _apply("times", _cn(0.02), _ci("x")), _apply("times", _cn(0.01), _ci("y"))))]
    rx = [_reaction("feed", [], [("S0", 1)], _ci("kf"))]
    for i in range(5):
        law = _apply("times", _ci("kc"), _ci(names[i])) if i != 2 else _apply("times", _ci("kc"), _ci(names[i]), _ci(names[i + 1]))
        rx.append(_reaction("c%d" % i, [(names[i], 1)], [(names[i + 1], 1)], law))
    rx.append(_reaction("deg", [("S5", 1)], [], _apply("times", _ci("kd"), _ci("S5"))))
    return _document("synthetic_chain6", axes, dims, species, params, assignments, rx)


def turing_dirichlet3d(nx, ny, nz):
    """The 3-D Turing model with Dirichlet values on both z sides of species_2 and zero fluxes everywhere else: the fused
    schedule (combine inside the last stage) together with the Dirichlet overwrite, incl. the reference's Z-side quirks.  dt 0.02."""
    bc = {("species_2", "Zmax"): (0.12, "Dirichlet"), ("species_2", "Zmin"): (0.09, "Dirichlet")}
    return turing(nx, ny, nz, bc_values=bc)


def advection_field(nx, ny, uniform=False, nz=None):
    """CIP-CSLR advection (cipCSLR, calcPDE.cpp:564-869) of two species in a box with velocities that are FIELDS — a shear flow
    u_x = 0.6 + 0.02 y, u_y = -0.4 + 0.015 x for P (positive and negative components) and a rotation-like field for Q — so
    that the face-averaged velocity (:612-613) and the divergence correction of the face values (:623-625) are live; a little
    diffusion and a decay reaction run beside it.  uniform=True: constant velocities (the large-grid advection benchmark).
    nz: the same in 3-D (a z component of either sign, the Z pass of the dimension split).  dt 0.05."""
    dims = [nx, ny] + ([nz] if nz else [])
    axes = ["x", "y", "z"][:len(dims)]
    kinds = {"x": "cartesianX", "y": "cartesianY", "z": "cartesianZ"}
    params = _transport_params("P", 0.05, axes) + _transport_params("Q", 0.02, axes)
    blob = [_apply("power", _apply("minus", _ci(a), _cn(f * (n - 1))), _cn(2)) for a, n, f in zip(axes, dims, (0.35, 0.55, 0.45))]
    box = []
    for a, n, (lo, hi) in zip(axes, dims, ((0.5, 0.7), (0.2, 0.45), (0.3, 0.8))):
        box += [_apply("gt", _ci(a), _cn(lo * (n - 1))), _apply("lt", _ci(a), _cn(hi * (n - 1)))]
    ia = [("P", _apply("plus", _cn(0.2), _apply("times", _cn(2.0), _apply("exp", _apply("minus", _apply("times", _cn(0.02), _apply("plus", *blob))))))),
          ("Q", _piecewise(_cn(3.0), _apply("and", *box), _cn(0.5)))]
    vel = {("P", "x"): (0.6, 0.02, "y"), ("P", "y"): (-0.4, 0.015, "x"), ("Q", "x"): (0.3, -0.02, "y"), ("Q", "y"): (-0.25, 0.02, "x")}
    if nz:
        vel.update({("P", "z"): (-0.5, 0.01, "x"), ("Q", "z"): (0.35, -0.015, "y")})
    for (sp_, a), (c0, c1, dep) in sorted(vel.items()):
        pid = "%s_adv_%s" % (sp_, a)
        inner = '<spatial:advectionCoefficient spatial:variable="%s" spatial:coordinate="%s"/>' % (sp_, kinds[a])
        if uniform:
            params.append(_param(pid, c0, inner))
        else:
            params.append('<parameter id="%s" constant="true">%s</parameter>' % (pid, inner))
            ia.append((pid, _apply("plus", _cn(c0), _apply("times", _cn(c1), _ci(dep)))))
    params.append(_param("kd", 0.01))
    rx = [_reaction("decay", [("P", 1)], [("Q", 1)], _apply("times", _ci("kd"), _ci("P")))]
    return _document("synthetic_advection_field", axes, dims, [("P", None), ("Q", None)], params, ia, rx)


def advection_uniform(nx, ny):
    return advection_field(nx, ny, uniform=True)


def advection_uniform3d(nx, ny, nz):
    return advection_field(nx, ny, uniform=True, nz=nz)


def advection_field3d(nx, ny, nz):
    return advection_field(nx, ny, uniform=False, nz=nz)
