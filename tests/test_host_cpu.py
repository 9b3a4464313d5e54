"""CPU tests (no GPU): the host half of the drop-in — SBML reading, geometry / membrane classification,
boundary types, reverse-polish streams, initial fields — against what the reference itself produced for the
same models (tests/golden/*.npz, written by tests/golden/make_golden.py from the unmodified reference
sources), plus the C ABI surface.

The bar for everything here is bit-exact (integer masks, index lists, token streams, initial values).
"""
import os
import re

import numpy as np
import pytest

from conftest import REPO, golden_cases, is_membrane_species, load_golden, model_path

import spatial_sbml_b200 as sb


def host_sim(case):
    g = load_golden(case)
    nx, ny, nz = [int(v) for v in g["dims"]]
    sim = sb.SpatialSimulator(host_only=True)
    sim.initFromFile(model_path(case), nx, ny, nz)
    return sim, g, (nx, ny, nz)


def even(a, dims):
    nx, ny, nz = dims
    return a.reshape(2 * nz - 1, 2 * ny - 1, 2 * nx - 1)[::2, ::2, ::2]


def test_library_exports_every_declared_symbol(lib):
    """include/*.h is the drop-in boundary: every SB2_API / SSB_API declaration must be exported."""
    names = []
    for h in ("spatial_b200.h", "spatial_sim_c.h"):
        text = open(os.path.join(REPO, "include", h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names += re.findall(r"S[SB][B2]_API\s+[\w\s\*]+?\b(\w+)\s*\(", text)
    assert len(names) > 60, names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, "declared in include/ but not exported: %s" % missing
    # and the ctypes prototypes of the Python binding cover exactly that surface
    assert sorted(lib._prototypes) == sorted(set(names))


def test_no_cpu_fallback():
    """Without a CUDA device a simulator that is not host-only must refuse to initialise (or, on a GPU box,
    initialise on the device): there is no CPU stepping path."""
    import torch
    sim = sb.SpatialSimulator()
    if torch.cuda.is_available():
        sim.initFromFile(model_path("turing_tiny"), 22, 22, 1)
        assert sim.getDeviceLaunchCount() == 0
    else:
        with pytest.raises(RuntimeError, match="CUDA"):
            sim.initFromFile(model_path("turing_tiny"), 22, 22, 1)
    sim.close()


def test_host_only_cannot_step():
    sim = sb.SpatialSimulator(host_only=True)
    sim.initFromFile(model_path("turing_tiny"), 22, 22, 1)
    with pytest.raises(RuntimeError):
        sim.oneStep(0, 0.07)
    sim.close()


@pytest.mark.parametrize("case", golden_cases())
def test_volume_masks_bit_exact(case):
    """isDomain, isBoundary and the six boundary-type flags of every volume compartment
    (reference spatialsimulator.cpp:411-717, boundaryFunction.cpp:13-217)."""
    sim, g, dims = host_sim(case)
    assert sim.getXDim() == dims[0] and sim.getYDim() == dims[1] and sim.getZDim() == dims[2]
    n_checked = 0
    for name in [str(s) for s in g["geo_names"]]:
        if not int(g["geo/%s/isVol" % name]):
            continue
        dom = sim.getGeometry(name)
        assert dom is not None, name
        assert dom.size == sim.getGeometryLength()
        assert np.array_equal(even(dom, dims).astype(np.int8), g["geo/%s/isDomain" % name]), "%s isDomain" % name
        # nothing may be set off the volume cells
        assert int(dom.sum()) == int(g["geo/%s/isDomain" % name].sum())
        bnd = sim.getBoundary(name)
        assert np.array_equal(even(bnd, dims).astype(np.int8), g["geo/%s/isBoundary" % name]), "%s isBoundary" % name
        bt = sim.getBoundaryType(name)
        bits = sum((bt[:, k].astype(np.uint8) << (k + 1)) for k in range(6))
        assert np.array_equal(even(bits, dims), g["geo/%s/bType" % name]), "%s bType" % name
        n_checked += 1
    assert n_checked >= 1
    sim.close()


@pytest.mark.parametrize("case", golden_cases())
def test_membrane_classification_bit_exact(case):
    """Membrane points (isDomain 1), pseudo-membrane corners (isDomain 2), their boundary types and the two index
    lists (reference spatialsimulator.cpp:721-1038)."""
    sim, g, dims = host_sim(case)
    for name in [str(s) for s in g["geo_names"]]:
        if int(g["geo/%s/isVol" % name]):
            continue
        dom = sim.getGeometry(name)
        assert dom is not None, name
        pts = np.flatnonzero(dom)
        assert np.array_equal(pts, g["geo/%s/points" % name]), "%s membrane points" % name
        assert np.array_equal(dom[pts].astype(np.int8), g["geo/%s/class" % name]), "%s membrane classes" % name
        bt = sim.getBoundaryType(name)
        bits = sum((bt[pts, k].astype(np.uint8) << (k + 1)) for k in range(6)).astype(np.uint8)
        assert np.array_equal(bits, g["geo/%s/bTypeAtPoints" % name]), "%s membrane bType" % name
        mem = np.array([int(v) for v in sim.getSetupText("memindex/" + name).split()], dtype=np.int64)
        assert np.array_equal(mem, g["geo/%s/domainIndex" % name]), "%s domainIndex" % name
        pseudo = np.array([int(v) for v in sim.getSetupText("pseudo/" + name).split()], dtype=np.int64)
        assert np.array_equal(pseudo, g["geo/%s/pseudoMemIndex" % name]), "%s pseudoMemIndex" % name
    sim.close()


@pytest.mark.parametrize("case", ["transport", "transport_rate", "syn_membrane_2d", "syn_membrane_box_2d", "syn_membrane_3d"])
def test_membrane_normals_bit_exact(case):
    """nuVec at the membrane points of the models whose transport reactions read it
    (reference setInfoFunction.cpp:495-1546)."""
    sim, g, dims = host_sim(case)
    checked = 0
    for name in [str(s) for s in g["geo_names"]]:
        key = "geo/%s/nuVec" % name
        if int(g["geo/%s/isVol" % name]) or key not in g.files:
            continue
        rows = [l.split() for l in sim.getSetupText("normals/" + name).splitlines()]
        if not rows:
            continue
        idx = np.array([int(r[0]) for r in rows])
        ours = np.array([[float(v) for v in r[1:4]] for r in rows])
        order = {int(v): k for k, v in enumerate(g["geo/%s/domainIndex" % name])}  # the dump lists nuVec in domainIndex order
        assert all(int(i) in order for i in idx), "%s: membrane point unknown to the reference" % name
        ref = g[key][[order[int(i)] for i in idx]]
        if not ours.any():
            continue  # normals are only built for membranes a transport reaction crosses
        assert np.array_equal(ours, ref), "%s normals" % name
        checked += 1
    assert checked >= 1
    sim.close()


@pytest.mark.parametrize("case", golden_cases())
def test_reverse_polish_streams_match_reference(case):
    """rearrangeAST + parseAST (reference astFunction.cpp:18-235): the token streams of every kinetic law and of every
    analytic-geometry expression, token for token, and the species references with stoichiometry / isVariable."""
    sim, g, dims = host_sim(case)
    i = 0
    while "rxn/%d/rpn" % i in g.files:
        assert sim.getSetupText("rxn/%d" % i) == str(g["rxn/%d/rpn" % i]), "reaction %d stream" % i
        assert sim.getSetupText("refs/%d" % i) == str(g["rxn/%d/refs" % i]), "reaction %d references" % i
        i += 1
    for name in [str(s) for s in g["geo_names"]]:
        key = "geo/%s/rpn" % name
        if key in g.files:
            assert sim.getSetupText("geo/" + name) == str(g[key]), "%s geometry stream" % name
    sim.close()


@pytest.mark.parametrize("case", golden_cases())
def test_initial_fields_bit_exact(case):
    """Initial concentrations incl. initial assignments (reference setInfoFunction.cpp:93-155,
    spatialsimulator.cpp:1051-1153) at the volume cells."""
    sim, g, dims = host_sim(case)
    for sp in [str(s) for s in g["species"]]:
        full = sim.getVariable(sp)
        assert full is not None, sp
        if is_membrane_species(g, sp):
            # a species on a membrane: its membrane and pseudo-membrane points (the latter filled with the smaller neighbour,
            # boundaryFunction.cpp:98-111), zero everywhere else
            pts = g["geo/%s/points" % str(g["memgeo/%s" % sp])]
            assert np.array_equal(np.asarray(full)[pts], g["step/0/" + sp]), "initial values of %s on its membrane" % sp
            assert not np.delete(np.asarray(full), pts).any()
            continue
        assert np.array_equal(even(full, dims), g["step/0/" + sp]), "initial field of %s" % sp
        got = sim.getVariableCompact(sp)
        assert np.array_equal(got, g["step/0/" + sp])
    sim.close()


@pytest.mark.parametrize("case", [c for c in golden_cases() if "membrane" in c])
def test_voronoi_neighbours_bit_exact(case):
    """setVoronoiInfo (reference setInfoFunction.cpp:1134-1546): the contour neighbours of every membrane point, the facet
    lengths s_ij (3-D: bisector intersections in the rotated tangent plane) and the distances d_ij after the symmetrising
    average, against the reference's own vorI array."""
    sim, g, dims = host_sim(case)
    checked = 0
    for name in [str(s) for s in g["geo_names"]]:
        key = "geo/%s/vorAdj" % name
        if key not in g.files or not len(g[key]):
            continue
        rows = np.array([[float(x) for x in l.split()] for l in sim.getSetupText("voronoi/" + name).strip().split("\n")])
        assert np.array_equal(rows[:, :6].astype(np.int64), g[key]), "%s neighbours" % name
        assert np.array_equal(rows[:, 6:12], g["geo/%s/vorSiDi" % name][:, :6]), "%s s_ij" % name
        assert np.array_equal(rows[:, 12:18], g["geo/%s/vorSiDi" % name][:, 6:]), "%s d_ij" % name
        checked += 1
    assert checked >= 1
    sim.close()


def test_get_variable_at_and_index_for_position():
    """getVariableAt takes coordinates, getIndexForPosition maps them to the doubled grid
    (reference spatialsimulator.cpp:1769-1783, :116)."""
    sim, g, dims = host_sim("turing_blob")
    u = g["step/0/species_1"]
    for (x, y) in [(0, 0), (10, 20), (50, 50), (100, 100)]:
        idx = sim.getIndexForPosition(x, y)
        assert idx == (2 * y) * (2 * dims[0] - 1) + 2 * x
        v = sim.getVariableAt("species_1", x, y, 0)
        dom = g["geo/%s/isDomain" % [str(s) for s in g["geo_names"] if int(g["geo/%s/isVol" % s])][0]][0, y, x]
        assert v == (u[0, y, x] if dom else 0.0)
    assert sim.getVariableAt("no_such_species", 1, 1, 0) == 0.0
    assert sim.getVariableLength() == sim.getGeometryLength() - 1  # the reference's N-1 (spatialsimulator.cpp:1341)
    sim.close()


def test_set_parameter_reaches_uniform_variables():
    sim, g, dims = host_sim("turing_tiny")
    text = sim.getSetupText("summary")
    assert "var " in text
    for name in ("k1", "Km", "species_2_diff"):  # kinetic constants and a diffusion coefficient (spatialsimulator.cpp:167-202)
        sim.setParameter(name, 123.25)
        line = [l for l in sim.getSetupText("summary").splitlines() if l.startswith("var %s " % name)][0]
        assert line.endswith("value 123.25"), line
    sim.setParameter("no_such_parameter", 1.0)  # unknown ids are ignored, as in the reference
    sim.close()


# ---------------------------------------------------------------------------------------------
# model-specialised stage kernel: the run-time build needs no GPU, only libnvrtc
# ---------------------------------------------------------------------------------------------
TURING_RECORDS = [("MULVC", 0, 0), ("MULV", 0, 1), ("XFER", 0, 1), ("ACCADDC1", 0, 0), ("ACCADDC1", 1, 0),
                  ("MULVC", 0, 1), ("ADDVC", 1, 1), ("RDIVC", 1, 0), ("MUL", 1, 0), ("ACCSUB1", 1, 0)]


def test_python_opcode_table_matches_the_device_header():
    text = open(os.path.join(REPO, "spatial_sbml_b200", "csrc", "device", "sb2_internal.h")).read()
    body = re.sub(r"//[^\n]*", "", re.search(r"enum Sb2Opc \{(.*?)\};", text, re.S).group(1))
    values, v = {}, -1
    for item in [i.strip() for i in body.split(",") if i.strip()]:
        if "=" in item:
            name, expr = [x.strip() for x in item.split("=")]
            v = values[expr] if expr in values else int(expr)
        else:
            name, v = item, v + 1
        values[name] = v
    for i, name in enumerate(sb.RD_OPCODES):
        assert values["OPC_" + name] == i, name
    assert values["OPC_RD_COUNT"] == len(sb.RD_OPCODES)
    assert "#define SB2_MAX_DEPTH %d " % sb.SB2_MAX_DEPTH in open(os.path.join(REPO, "include", "spatial_b200.h")).read()


def _nvrtc_present():
    import ctypes
    for name in (os.environ.get("SB2_NVRTC_LIB"), "libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12"):
        if not name:
            continue
        try:
            ctypes.CDLL(name)
            return True
        except OSError:
            pass
    return False


@pytest.mark.skipif(not _nvrtc_present(), reason="libnvrtc.so.12 not installed")
@pytest.mark.parametrize("d3", [False, True])
def test_specialised_kernel_builds_without_a_gpu(tmp_path, d3):
    """The Turing law (as the lowering pass emits it) expands to straight-line code that NVRTC compiles for sm_100a."""
    cubin = str(tmp_path / "turing.cubin")
    rc, log = sb.specialize_dry_run(TURING_RECORDS, ns=2, np_=2, w=8, depth=2, minb=2, d3=d3, cubin_path=cubin)
    assert rc == 0, log
    assert "error" not in log.lower(), log
    data = open(cubin, "rb").read()
    assert data[:4] == b"\x7fELF"
    for name in (b"rd_jit_first", b"rd_jit_middle", b"rd_jit_final"):
        assert name in data


@pytest.mark.skipif(not _nvrtc_present(), reason="libnvrtc.so.12 not installed")
def test_specialised_build_covers_every_opcode():
    """One record of every instruction form (stack slots chosen so that the program is well formed) compiles."""
    recs = [("PUSHC", 0, 0), ("PUSHV", 1, 0), ("PUSHF", 2, 3)]
    for op in ("ADD", "SUB", "MUL", "DIV", "GEQ", "GT", "LEQ", "LT", "AND", "OR"):
        recs += [("PUSHV", 2, 1), (op, 2, 0)]
    for op in ("ADDC", "SUBC", "MULC", "DIVC", "RSUBC", "RDIVC", "GEQC", "GTC", "LEQC", "LTC"):
        recs.append((op, 1, 0))
    for op in ("ADDV", "SUBV", "MULV", "DIVV"):
        recs.append((op, 1, 1))
    for op in ("ADDVC", "SUBVC", "RSUBVC", "MULVC", "DIVVC", "RDIVVC"):
        recs.append((op, 2, 1))
    for op in ("ADDVV", "SUBVV", "MULVV", "DIVVV"):
        recs.append((op, 2, 0 | (1 << 8)))
    for fn in (sb.OPS["EXP"], sb.OPS["NOT"], sb.OPS["ROOT"]):
        recs.append(("UNARY", 2, fn))
    for fn in (sb.OPS["POWER"], sb.OPS["LOG"], sb.OPS["XOR"], sb.OPS["EQ"], sb.OPS["NEQ"]):
        recs += [("PUSHV", 3, 0), ("BINR", 3, fn)]
    recs += [("MOVUP", 2, 0), ("MOV0", 3, 0), ("MASK", 0, 1)]
    for op in ("ACCSUB", "ACCADD", "ACCSUB1", "ACCADD1", "ACCSUBC1", "ACCADDC1", "ACCSUBM", "ACCADDM"):
        recs.append((op, 1, 0))
    recs.append(("XFER", 0, 2))
    assert {r[0] for r in recs} == set(sb.RD_OPCODES) - {"END"}
    rc, log = sb.specialize_dry_run(recs, ns=3, np_=2, w=4, depth=4)
    assert rc == 0, log


@pytest.mark.skipif(not _nvrtc_present(), reason="libnvrtc.so.12 not installed")
def test_specialised_build_of_an_empty_and_of_a_maximal_program():
    """A model without reactions (pure diffusion) and a program of SB2_RD_MAX_TOKENS records (3 KB of constants in the
    kernel parameters on top of the stage arguments) both build."""
    rc, log = sb.specialize_dry_run([], ns=1, np_=2, w=8, depth=1)
    assert rc == 0, log
    recs = [("PUSHV", 0, 0)] + [("ADDC", 0, 0)] * 382 + [("ACCADD1", 0, 0)]
    assert len(recs) == 384
    rc, log = sb.specialize_dry_run(recs, ns=1, np_=1, w=8, depth=1, minb=3, d3=True)
    assert rc == 0, log


@pytest.mark.skipif(not _nvrtc_present(), reason="libnvrtc.so.12 not installed")
@pytest.mark.parametrize("ns,np_,w,minb,d3", [(1, 2, 8, 2, False), (3, 2, 4, 2, False), (4, 1, 4, 2, False), (3, 1, 8, 2, True),
                                              (4, 1, 8, 2, True), (2, 2, 8, 2, True), (2, 4, 4, 2, True)])
def test_specialised_build_for_every_kernel_shape(ns, np_, w, minb, d3):
    """Species counts and strip / CTA shapes the library can ask for (2-D row marching and 3-D plane marching): the
    cascade law (a division of two registers, a comparison, stoichiometry 2) builds for each."""
    last = ns - 1
    recs = [("PUSHV", 0, 0), ("MULC", 0, 0), ("ACCSUB1", 0, 0), ("ACCADD", last, 0),
            ("PUSHV", 0, last), ("MULC", 0, 0), ("PUSHV", 1, last), ("ADDC", 1, 0), ("DIV", 1, 0),
            ("PUSHV", 1, last), ("GTC", 1, 0), ("MUL", 1, 0), ("ACCSUB1", last, 0)]
    rc, log = sb.specialize_dry_run(recs, ns=ns, np_=np_, w=w, depth=2, minb=minb, d3=d3)
    assert rc == 0, log
    assert "error" not in log.lower(), log


def test_specialised_build_rejects_oversized_programs():
    rc, log = sb.specialize_dry_run([("ADDC", 0, 0)] * 385, ns=2, np_=2, w=8, depth=1)  # SB2_RD_MAX_TOKENS = 384
    assert rc == -1


def test_threaded_host_setup_gives_the_same_model(monkeypatch):
    """Grids of 2^18 cells and more evaluate the geometry / initial-assignment streams and the boundary scan on several
    host threads (SSB_HOST_THREADS, default: all cores): masks, boundary types and initial fields must not depend on it."""
    from spatial_sbml_b200 import synthetic
    dims = (640, 520, 1)
    model = synthetic.turing(dims[0], dims[1])
    out = {}
    for threads in ("1", "7"):
        monkeypatch.setenv("SSB_HOST_THREADS", threads)
        sim = sb.SpatialSimulator(host_only=True)
        sim.initFromString(model, *dims)
        out[threads] = {"geo": sim.getGeometry("inside").copy(), "bnd": sim.getBoundary("inside").copy(),
                        "btype": sim.getBoundaryType("inside").copy(), "u": sim.getVariable("species_1").copy(),
                        "v": sim.getVariable("species_2").copy(), "summary": sim.getSetupText("summary")}
        sim.close()
    for key in ("geo", "bnd", "btype", "u", "v"):
        assert np.array_equal(out["1"][key], out["7"][key]), key
    assert out["1"]["summary"] == out["7"]["summary"]
    assert out["1"]["geo"].sum() == (dims[0] - 2) * (dims[1] - 2)


def test_dmp_round_trip_and_clamping(tmp_path):
    """DataHelper (datahelper.hh): "maxX maxY" + maxX lines of maxY numbers in default iostream formatting, data[x][y],
    operator() floors and clamps"""
    from spatial_sbml_b200.dmp import DataHelper
    h = DataHelper(rows=3, cols=2)
    h.data[:] = [[0.5, 1234567.0], [1e-7, -2.25], [3.0, 0.1 + 0.2]]
    f = str(tmp_path / "a.dmp")
    h.write(f)
    assert open(f).read() == "3 2\n0.5 1.23457e+06 \n1e-07 -2.25 \n3 0.3 \n"
    g = DataHelper(f)
    assert g.is_valid() and g.data.shape == (3, 2)
    assert g(1.9, 1.2) == -2.25 and g(-4, 7) == 1234570.0 and g(99, 0) == 3.0
    assert not DataHelper(str(tmp_path / "missing.dmp")).is_valid()


def test_dmp_import_export_follow_the_ui(tmp_path):
    """exportConcentration collects getVariableAt(x, y); importConcentration doses every cell of the species' compartment and
    leaves the others alone (spatialmainwidget.cpp:205-257, simulationthread.cpp:160-185)."""
    from spatial_sbml_b200.dmp import DataHelper
    sim, g, dims = host_sim("turing_tiny")
    nx, ny = dims[0], dims[1]
    h = DataHelper.from_simulator(sim, "species_1")
    assert h.data.shape == (nx, ny)
    for x, y in ((0, 0), (5, 7), (nx - 1, 3), (10, ny - 2)):
        assert h(x, y) == sim.getVariableAt("species_1", x, y, 0)
    f = str(tmp_path / "s.dmp")
    h.data[:] = np.arange(nx * ny, dtype=np.float64).reshape(nx, ny) / 8.0  # exactly representable with six digits
    h.write(f)
    before = sim.getVariableCompact("species_1").copy()
    assert sim.getSpeciesCompartment("species_1") is not None
    assert DataHelper(f).apply(sim, "species_1")
    after = sim.getVariableCompact("species_1")
    geo = np.asarray(sim.getGeometry(sim.getSpeciesCompartment("species_1"))).reshape(2 * ny - 1, 2 * nx - 1)[::2, ::2] != 0
    assert geo.any() and not geo.all()
    assert np.array_equal(after[0][geo], (h.data.T)[geo])
    assert np.array_equal(after[0][~geo], before[0][~geo])
    sim.close()


@pytest.mark.parametrize("case", ["syn_membrane_2d", "syn_membrane_3d"])
def test_membrane_csv_of_the_cli_lists_every_doubled_grid_point(case, tmp_path):
    """outputTimeCource's membrane file (outputFunction.cpp:65-98, :133-152, :175-198): one row per point of the doubled grid with
    `mFlag, value` per membrane species, a blank line after each row of points (3-D: not after the last row of a plane)."""
    from spatial_sbml_b200 import cli
    sim, g, dims = host_sim(case)
    nx, ny, nz = dims
    X, Y, Z = 2 * nx - 1, 2 * ny - 1, 2 * nz - 1
    mem = cli.membrane_species(sim)
    assert mem and all(is_membrane_species(g, sid) for sid, _ in mem)
    vol = [n for n in cli.spatial_species(sim) if n not in set(s for s, _ in mem)]
    assert vol and not any(is_membrane_species(g, n) for n in vol)
    f = str(tmp_path / "m.csv")
    cli.write_membrane_csv(f, sim, mem, case, 1.0, 0.1, 0.0)
    lines = open(f).read().split("\n")
    assert lines[0] == "# results of %s.xml simulation" % case
    assert lines[2].startswith("# about mFlag")
    assert lines[5] == "# x, y%s" % (", z" if nz > 1 else "") + "".join(", mFlag, " + sid for sid, _ in mem)
    body = lines[6:]
    rows = [l for l in body if l]
    assert len(rows) == X * Y * Z
    blanks = len([l for l in body[:-1] if not l])
    assert blanks == (Y if nz == 1 else Z * (Y - 1))
    t = np.array([[float(v) for v in r.split(", ")] for r in rows])
    c0 = 3 if nz > 1 else 2
    sid, comp = mem[0]
    flags = t[:, c0].astype(int)
    assert set(np.unique(flags)) == {0, 1, 2}
    assert np.array_equal(flags, np.asarray(sim.getGeometry(comp))[:X * Y * Z])
    pts = g["geo/%s/points" % str(g["memgeo/%s" % sid])]
    assert np.array_equal(t[pts, c0 + 1], g["step/0/" + sid])
    assert not t[flags == 0, c0 + 1].any()
    sim.close()


def test_the_references_own_cli_compiles_against_our_header():
    """Source compatibility of the drop-in class: SpatialCL/main.cpp of the reference, unmodified and in place, compiled against
    csrc/host/spatialsimulator.h and linked with libspatialb200.so (oracle/Makefile, target spatialcl_b200).  Without a GPU the
    program must fail loudly - there is no CPU path to fall back to."""
    import subprocess
    exe = os.path.join(REPO, "oracle", "_ref", "spatialcl_b200")
    if os.path.isdir("/root/reference/SpatialCL"):
        subprocess.run(["make", "-s", "-C", os.path.join(REPO, "oracle"), exe], check=True)
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/spatialcl_b200 is built where /root/reference is mounted")
    import torch
    if torch.cuda.is_available():
        pytest.skip("run on the GPU by tests/test_parity_gpu.py")
    r = subprocess.run([exe, model_path("turing_tiny")], capture_output=True, text=True)
    assert r.returncode != 0
    assert "spatial required: yes" in r.stdout
    assert "no CPU fallback" in r.stderr


def test_the_bridge_of_the_integration_guide_builds():
    """INTEGRATION.md §A, compiled: oracle/ref_bridge.cpp links the reference's own unmodified objects (its host set-up) with
    libspatialb200.so and drives the sb2_* C ABI from the reference's structs.  Without a GPU it must stop at sb2_create."""
    import subprocess
    exe = os.path.join(REPO, "oracle", "_ref", "ref_bridge")
    if os.path.isdir("/root/reference/SpatialSBML"):
        subprocess.run(["make", "-s", "-C", os.path.join(REPO, "oracle"), exe], check=True)
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_bridge is built where /root/reference is mounted")
    import torch
    if torch.cuda.is_available():
        pytest.skip("run on the GPU by tests/test_parity_gpu.py")
    r = subprocess.run([exe, model_path("turing_tiny"), "22", "22", "1", "0.07", "2"], capture_output=True, text=True)
    assert r.returncode == 2 and "no CPU fallback" in r.stderr
