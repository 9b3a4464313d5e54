"""GPU tests of the slab-sharded step: sharded == single device, bit for bit (no reduction inside the step).

The sharded step is orchestrated inside the library (sb2_sharded_begin / sb2_sharded_next); SpatialSimulator(devices=[...])
steps one slab per listed device with peer copies of the halo rows (sb2_group_step).  A device ordinal may repeat, so the slab
logic — edge rows first, the split boundary pass, halo exchanges per stage / per advection pass / after the update, the
two-row halos of advection — runs on a box with ONE GPU.  With two or more GPUs the torchrun + NCCL path is run as well.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import REPO, load_golden, model_path

pytestmark = pytest.mark.gpu


def _run_pair(model, dims, dt, steps, species, devices, jit=None, monkeypatch=None, from_file=False):
    import spatial_sbml_b200 as sb
    if jit and monkeypatch is not None:
        monkeypatch.setenv("SB2_JIT", jit)
    sims = []
    for devs in (None, devices):
        s = sb.SpatialSimulator(devices=devs)
        if from_file:
            s.initFromFile(model, *dims)
        else:
            s.initFromString(model, *dims)
        sims.append(s)
    one, many = sims
    assert many.getNumSlabs() == len(devices)
    for t in range(steps):
        one.oneStep(t, dt)
        many.oneStep(t, dt)
    for sp in species:
        a, b = one.getVariableCompact(sp), many.getVariableCompact(sp)
        assert np.isfinite(a).all()
        assert np.array_equal(a, b), "%s differs between 1 device and %d slabs: %d cells, max abs %.3e" % (
            sp, len(devices), int((a != b).sum()), float(np.abs(a - b).max()))
    out = [one.getVariableCompact(sp) for sp in species]
    one.close()
    many.close()
    return out


@pytest.mark.parametrize("nslabs", [2, 3])
@pytest.mark.parametrize("d3", [False, True])
def test_slabs_on_the_other_sharded_schedule(nslabs, d3, monkeypatch):
    """Slabs of eight rows or more exchange once per step and recompute the apron rows (sb2_sharded_set_apron; the tests below);
    SB2_NO_APRON keeps the schedule with an exchange after every stage, which thin slabs and unfused models always take."""
    from spatial_sbml_b200 import synthetic
    # 2-D grids take the apron schedule by default, 3-D grids the exchange per stage: run the other one here
    monkeypatch.setenv("SB2_APRON" if d3 else "SB2_NO_APRON", "1")
    dims = (140, 24, 9 * nslabs + 2) if d3 else (300, 50 * nslabs + 7, 1)
    model = synthetic.turing(*dims) if d3 else synthetic.turing(dims[0], dims[1])
    _run_pair(model, dims, 0.02 if d3 else 0.07, 8, ("species_1", "species_2"), [0] * nslabs)


@pytest.mark.parametrize("nslabs", [2, 3])
@pytest.mark.parametrize("jit", [None, "1"])
def test_slabs_on_one_device_turing_2d(nslabs, jit, monkeypatch):
    from spatial_sbml_b200 import synthetic
    dims = (300, 50 * nslabs + 7, 1)
    u = _run_pair(synthetic.turing(dims[0], dims[1]), dims, 0.07, 10, ("species_1", "species_2"), [0] * nslabs, jit, monkeypatch)
    assert u[0].std() > 0


@pytest.mark.parametrize("nslabs", [2, 3])
@pytest.mark.parametrize("jit", [None, "1"])
def test_slabs_on_one_device_turing_3d(nslabs, jit, monkeypatch):
    from spatial_sbml_b200 import synthetic
    dims = (140, 24, 9 * nslabs + 2)
    _run_pair(synthetic.turing(*dims), dims, 0.02, 8, ("species_1", "species_2"), [0] * nslabs, jit, monkeypatch)


def test_thin_slabs_whose_halo_rows_cover_the_whole_grid():
    """Slabs of three rows hold, with their four halo rows either side, the whole grid: they are slabs all the same."""
    from spatial_sbml_b200 import synthetic
    dims = (96, 10, 1)
    _run_pair(synthetic.turing_bc(dims[0], dims[1]), dims, 0.05, 8, ("species_1", "species_2"), [0, 0, 0])


@pytest.mark.parametrize("nslabs", [2, 3])
def test_slabs_with_live_boundary_conditions(nslabs):
    """Non-zero Neumann fluxes (carried in k0 between steps) and a Dirichlet side: the boundary pass is split into edge rows and
    interior rows like the stage kernels, Dirichlet values travel with the stage exchanges, the new u after the update."""
    from spatial_sbml_b200 import synthetic
    dims = (96, 20 * nslabs + 3, 1)
    _run_pair(synthetic.turing_bc(dims[0], dims[1]), dims, 0.05, 12, ("species_1", "species_2"), [0] * nslabs)


@pytest.mark.parametrize("case,nslabs", [("fish_dirichlet", 2), ("fish_dirichlet", 3), ("turing_tiny_flux", 2), ("syn_turing_3d_flux", 2),
                                         ("syn_turing_3d_flux", 3), ("syn_turing_3d_dirichlet", 2), ("syn_turing_3d_dirichlet", 3)])
def test_slabs_golden_models_with_boundary_conditions(case, nslabs):
    g = load_golden(case)
    dims = tuple(int(v) for v in g["dims"])
    species = [str(s) for s in g["species"]]
    out = _run_pair(model_path(case), dims, float(g["dt"]), 10, species, [0] * nslabs, from_file=True)
    for sp, got in zip(species, out):
        ref = g["step/10/" + sp]
        assert np.abs(got - ref).max() <= 1e-10 * max(np.abs(ref).max(), 1e-300), sp


@pytest.mark.parametrize("case", ["syn_advection_uniform_3d", "syn_advection_field_3d", "syn_advection_field_2d"])
@pytest.mark.parametrize("nslabs", [2, 3])
def test_slabs_with_advection_more_models(case, nslabs):
    """z-slabs of a 3-D grid (three direction passes per species, two halo planes of u and of the three face arrays) and
    field-valued velocities on slabs"""
    g = load_golden(case)
    dims = tuple(int(v) for v in g["dims"])
    species = [str(s) for s in g["species"]]
    out = _run_pair(model_path(case), dims, float(g["dt"]), 10, species, [0] * nslabs, from_file=True)
    for sp, got in zip(species, out):
        ref = g["step/10/" + sp]
        assert np.abs(got - ref).max() <= 1e-10 * max(np.abs(ref).max(), 1e-300), sp


@pytest.mark.parametrize("nslabs", [2, 3, 4])
def test_slabs_with_advection(nslabs):
    """CIP-CSLR advection on slabs: two halo rows of u and of the face values, exchanged after every direction pass; the golden
    advection model (three species, velocities of both signs on both axes) against the reference's own fields too."""
    g = load_golden("advection")
    dims = tuple(int(v) for v in g["dims"])
    species = [str(s) for s in g["species"]]
    out = _run_pair(model_path("advection"), dims, float(g["dt"]), 10, species, [0] * nslabs, from_file=True)
    for sp, got in zip(species, out):
        ref = g["step/10/" + sp]
        assert np.abs(got - ref).max() <= 1e-10 * max(np.abs(ref).max(), 1e-300), sp


@pytest.mark.parametrize("case", ["transport", "transport_rate", "syn_transport_mixed_2d", "syn_sampled_2d"])
@pytest.mark.parametrize("nslabs", [2, 3, 5])
def test_slabs_with_membrane_transport(case, nslabs):
    """calcMemTransport on slabs: every slab evaluates the membrane points next to its rows (taken from a host model of the whole
    grid: the reference's frame tests and normals refer to the whole doubled grid), reads operands up to two rows into its halo
    (k rows travel two deep per stage) and accumulates into the cells it owns."""
    g = load_golden(case)
    dims = tuple(int(v) for v in g["dims"])
    species = [str(s) for s in g["species"]]
    out = _run_pair(model_path(case), dims, float(g["dt"]), 10, species, [0] * nslabs, from_file=True)
    for sp, got in zip(species, out):
        ref = g["step/10/" + sp]
        assert np.abs(got - ref).max() <= 1e-10 * max(np.abs(ref).max(), 1e-300), sp


def test_group_accessors():
    """setVariableCompact / getVariableCompact / getVariable / getVariableAt / setParameter / getStableTimeStep on a grid held
    by several slabs behave like on one device."""
    import spatial_sbml_b200 as sb
    from spatial_sbml_b200 import synthetic
    dims = (64, 45, 1)
    model = synthetic.brusselator(dims[0], dims[1], perturb=True)
    a = sb.SpatialSimulator()
    a.initFromString(model, *dims)
    b = sb.SpatialSimulator(devices=[0, 0, 0])
    b.initFromString(model, *dims)
    x0 = a.getVariableCompact("X").copy()
    x0[0, 10:30, 5:40] += 0.25
    for s in (a, b):
        s.setVariableCompact("X", x0)
        s.setParameter("k", 0.12)
        s.runSteps(0, 0.1, 6)
    assert np.array_equal(a.getVariableCompact("X"), b.getVariableCompact("X"))
    assert np.array_equal(a.getVariableCompact("Y"), b.getVariableCompact("Y"))
    assert np.array_equal(np.asarray(a.getVariable("X")), np.asarray(b.getVariable("X")))
    assert a.getVariableAt("X", 20, 30, 0) == b.getVariableAt("X", 20, 30, 0)
    for dt in (0.1, 10.0):
        assert a.getStableTimeStep(dt) == b.getStableTimeStep(dt)
    assert a.getNumDomainCells() == b.getNumDomainCells()
    a.close()
    b.close()


@pytest.mark.parametrize("nslabs", [2, 3])
@pytest.mark.parametrize("d3", [False, True])
@pytest.mark.parametrize("jit", [None, "1"])
def test_group_host_step_is_the_same_step(nslabs, d3, jit, monkeypatch):
    """oneStepHost on a grid held by several slabs (sb2_group_step_host: every slab pipelines upload / stages / download over its
    own rows, ONE exchange of four edge rows of u per step, apron rows recomputed) gives the fields of oneStep on one device, also
    when plain steps (which must first refresh the halo rows a host step leaves unexchanged) are mixed in."""
    import spatial_sbml_b200 as sb
    from spatial_sbml_b200 import synthetic
    if jit:
        monkeypatch.setenv("SB2_JIT", jit)
    dims = (140, 24, 17 * nslabs + 3) if d3 else (300, 41 * nslabs + 5, 1)
    dt = 0.02 if d3 else 0.07
    model = synthetic.turing(*dims) if d3 else synthetic.turing(dims[0], dims[1])
    species = ("species_1", "species_2")
    one = sb.SpatialSimulator()
    one.initFromString(model, *dims)
    many = sb.SpatialSimulator(devices=[0] * nslabs)
    many.initFromString(model, *dims)
    cur = {sp: many.getVariableCompact(sp).copy() for sp in species}
    for t in range(7):
        one.oneStep(t, dt)
        if t == 3:
            many.oneStep(t, dt)
            cur = {sp: many.getVariableCompact(sp).copy() for sp in species}
        else:
            nxt = {sp: np.empty_like(cur[sp]) for sp in species}
            many.oneStepHost(t, dt, cur, nxt, 2)
            cur = nxt
        for sp in species:
            want = one.getVariableCompact(sp)
            assert np.array_equal(cur[sp], want), (t, sp, int((cur[sp] != want).sum()))
    for sp in species:
        assert np.array_equal(many.getVariableCompact(sp), cur[sp])
    assert many.getDeviceLaunchCount() > 0
    one.close()
    many.close()


def test_group_host_step_falls_back_for_unfused_models():
    import spatial_sbml_b200 as sb
    g = load_golden("advection")
    dims = tuple(int(v) for v in g["dims"])
    species = [str(s) for s in g["species"]]
    many = sb.SpatialSimulator(devices=[0, 0])
    many.initFromFile(model_path("advection"), *dims)
    cur = {sp: many.getVariableCompact(sp).copy() for sp in species}
    out = {sp: np.empty_like(cur[sp]) for sp in species}
    many.oneStepHost(0, float(g["dt"]), cur, out)
    for sp in species:
        ref = g["step/1/" + sp]
        assert np.abs(out[sp] - ref).max() <= 1e-10 * np.abs(ref).max(), sp
    many.close()


@pytest.mark.parametrize("dim", ["2", "3", "bc", "adv", "mt", "host2", "host3"])
def test_nccl_sharded_step_is_bit_identical(dim):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("the NCCL path needs at least two GPUs (the same slab logic runs on one GPU in the tests above)")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + ["2", "3", "bc", "adv", "mt", "host2", "host3"].index(dim)), os.path.join(REPO, "tools", "sharded_check.py"), dim, "12"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "BIT-IDENTICAL" in r.stdout
