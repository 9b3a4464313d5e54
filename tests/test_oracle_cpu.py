"""CPU tests of the checkers themselves (no GPU).

  * oracle/rd_oracle.py — the numpy restatement of the reference's reaction-diffusion-advection step — against the fields the
    UNMODIFIED reference produced (tests/golden/*.npz): bit for bit, every stored step.
  * oracle/_ref/ref_driver — the reference's own sources compiled in place — re-run here on a small model must
    reproduce the committed golden exactly (skipped where the binary has not been built).
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import REPO, load_golden, model_path

sys.path.insert(0, os.path.join(REPO, "oracle"))
import rd_oracle  # noqa: E402
from sbd import read_sbd  # noqa: E402

import spatial_sbml_b200 as sb

# case -> last stored step the restatement is run to (all stored steps up to it are compared)
RESTATED = {"turing_tiny": 1000, "turing_tiny_cl": 100, "turing_uniform": 100, "turing_blob": 100, "turing_50": 100,
            "brusselator": 100, "brusselator_nonunit": 100, "brusselator_300": 1, "fish": 10, "fish_300": 1,
            "syn_turing_2d": 100, "syn_turing_3d": 50, "syn_brusselator_3d": 20, "syn_cascade3_3d": 40, "syn_cascade3_2d": 40,
            # calcBoundary restated too: non-zero Neumann flux (carried in k0 between steps) and a Dirichlet side
            "turing_tiny_flux": 100, "fish_dirichlet": 10, "syn_turing_3d_flux": 40, "syn_turing_3d_dirichlet": 40,
            # cipCSLR restated on the doubled grid (cell + face values)
            "advection": 100,
            # calcMemTransport restated per membrane point (incl. the rpStack[1] rate quirk)
            "transport": 100, "transport_rate": 100,
            # ... with ordinary reactions on the transported species before and after the transport in the reaction list
            "syn_transport_mixed_2d": 100,
            # rate rules only; power, log to a base, xor / eq / neq, tanh, cos, time and a pop-only operator in the laws
            "syn_raterules_2d": 100,
            # geometry from an image + a membrane transport; six species
            "syn_sampled_2d": 100, "syn_chain6_2d": 100,
            "syn_advection_uniform_2d": 100, "syn_advection_uniform_3d": 40,
            # field-valued advection velocities (face-averaged velocity, divergence correction) and diffusion coefficients
            "syn_advection_field_2d": 100, "syn_advection_field_3d": 40, "syn_fieldcoeff_2d": 100, "syn_fieldcoeff_3d": 40,
            # assignment rules re-evaluated every step (updateAssignmentRules) on parameters and on a species, field operands in laws
            "syn_rules_2d": 100, "syn_rules_3d": 40,
            # species ON a membrane: calcMemDiffusion over the Voronoi neighbours, reactions on the membrane, binding / unbinding
            # through calcMemTransport, the pseudo-membrane fill (round / box interface in 2-D, a sphere in 3-D, mixed with
            # ordinary reactions)
            "syn_membrane_2d": 100, "syn_membrane_box_2d": 100, "syn_membrane_3d": 50, "syn_membrane_mixed_2d": 100}


def build(case):
    g = load_golden(case)
    dims = tuple(int(v) for v in g["dims"])
    sim = sb.SpatialSimulator(host_only=True)
    sim.initFromFile(model_path(case), *dims)
    orc = rd_oracle.from_simulator(sim, open(model_path(case)).read(), dims)
    sim.close()
    return orc, g


@pytest.mark.parametrize("case", sorted(RESTATED))
def test_restatement_matches_reference_bit_for_bit(case):
    orc, g = build(case)
    dt = float(g["dt"])
    done = 0
    for target in [int(s) for s in g["steps"] if int(s) <= RESTATED[case]]:
        for t in range(done, target):
            assert orc.one_step(t, dt) == t + dt
        done = target
        for sp in [str(s) for s in g["species"]]:
            ref = g["step/%d/%s" % (target, sp)]
            # a species on a membrane is stored at the membrane's points (membrane + pseudo-membrane, ascending index)
            got = orc.sp[sp].u if sp in orc.sp else orc.msp[sp].val[g["geo/%s/points" % orc.msp[sp].comp]]
            assert np.array_equal(got, ref), "%s step %d species %s: max abs diff %.3e" % (case, target, sp, np.abs(got - ref).max())
    assert done > 0


# Goldens the numpy restatement does not cover: none.
NOT_RESTATED = set()


def test_every_golden_case_is_restated():
    """The restatement covers every path the goldens exercise except the ones listed (with the reason) above."""
    from conftest import golden_cases
    assert sorted(set(RESTATED) | NOT_RESTATED) == sorted(golden_cases())
    assert not (set(RESTATED) & NOT_RESTATED)


@pytest.mark.parametrize("case", ["turing_tiny", "syn_turing_3d"])
def test_reference_build_reproduces_the_golden(case, tmp_path):
    drv = os.path.join(REPO, "oracle", "_ref", "ref_driver")
    if not os.path.exists(drv):
        pytest.skip("oracle/_ref/ref_driver is not built (needs /root/reference; __graft_entry__.build() makes it)")
    g = load_golden(case)
    nx, ny, nz = [int(v) for v in g["dims"]]
    steps = [int(s) for s in g["steps"]][:2]
    out = str(tmp_path / "dump.sbd")
    subprocess.run([drv, "dump", model_path(case), str(nx), str(ny), str(nz), repr(float(g["dt"])), out] + [str(s) for s in steps],
                   check=True, stdout=subprocess.DEVNULL)
    d = read_sbd(out)
    for s in [0] + steps:
        for sp in [str(v) for v in g["species"]]:
            full = d["step/%d/%s" % (s, sp)].reshape(2 * nz - 1, 2 * ny - 1, 2 * nx - 1)
            assert np.array_equal(full[::2, ::2, ::2], g["step/%d/%s" % (s, sp)]), "%s step %d %s" % (case, s, sp)


@pytest.mark.parametrize("case", ["turing_tiny", "syn_advection_field_2d", "syn_advection_field_3d", "syn_fieldcoeff_2d", "fish"])
def test_every_variable_array_equals_the_references(case, tmp_path):
    """getVariable() of EVERY non-uniform variable - species, coordinates, parameters set by initial assignments (which the
    reference evaluates on every point of the doubled grid, spatialsimulator.cpp:1126-1139) - against the reference's own value[]
    arrays, dumped by the reference build for this test."""
    drv = os.path.join(REPO, "oracle", "_ref", "ref_driver")
    if not os.path.exists(drv):
        pytest.skip("oracle/_ref/ref_driver is not built (needs /root/reference; __graft_entry__.build() makes it)")
    g = load_golden(case)
    nx, ny, nz = [int(v) for v in g["dims"]]
    out = str(tmp_path / "dump.sbd")
    subprocess.run([drv, "dump", model_path(case), str(nx), str(ny), str(nz), repr(float(g["dt"])), out], check=True, stdout=subprocess.DEVNULL)
    d = read_sbd(out)
    sim = sb.SpatialSimulator(host_only=True)
    sim.initFromFile(model_path(case), nx, ny, nz)
    checked = 0
    for name in d["varNames"].split():
        key = "var/%s/value" % name
        if key not in d or d[key].size <= 1:
            continue
        ours = sim.getVariable(name)
        assert ours is not None, name
        assert np.array_equal(np.asarray(ours)[:d[key].size], d[key]), name
        checked += 1
    assert checked >= 4
    sim.close()
