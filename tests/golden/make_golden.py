#!/usr/bin/env python
"""Generates the golden fixtures under tests/golden/ by RUNNING THE REFERENCE ITSELF.

The reference (fbergmann/spatial-sbml) ships no tests and no golden vectors, so the
fixtures are outputs of its own, unmodified SpatialSBML sources, compiled in place from
/root/reference by oracle/Makefile (→ oracle/_ref/ref_driver).  This script only runs in the build
container (it reads /root/reference); the GPU box uses the committed results.

For every case it writes
  tests/golden/models/<case>.xml   the model the oracle was run on: the reference's example model
                                   with annotations / unit definitions / notes stripped (the simulator
                                   reads none of them), optionally with a few parameter values changed,
                                   or a synthetic model from spatial_sbml_b200/synthetic.py
  tests/golden/<case>.npz          dims, dt, geometry masks (isDomain, isBoundary at volume cells,
                                   bType, membrane domainIndex / pseudoMemIndex, normals), the
                                   reference's reverse-polish streams, and the species fields at the
                                   volume cells (plus the odd "face" points for advected species)
                                   after the listed numbers of oneStep calls.

usage: python tests/golden/make_golden.py [case ...]
"""
import os
import re
import subprocess
import sys
import xml.etree.ElementTree as ET

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "oracle"))
from sbd import read_sbd  # noqa: E402
from spatial_sbml_b200 import synthetic  # noqa: E402

EXAMPLES = "/root/reference/examples"
DRIVER = os.path.join(REPO, "oracle", "_ref", "ref_driver")
CORE = "http://www.sbml.org/sbml/level3/version1/core"
SPATIAL = "http://www.sbml.org/sbml/level3/version1/spatial/version1"
MATHML = "http://www.w3.org/1998/Math/MathML"

# name: (source, dims, dt, steps, parameter overrides, advected species whose face points are stored)
CASES = {
    # BASELINE config 1 (UI convention) and the CL convention of SpatialCL/main.cpp:53-56
    "turing_tiny": ("turing_parameters_tiny.xml", (22, 22, 1), 0.07, [1, 10, 100, 1000], {}, []),
    "turing_tiny_cl": ("turing_parameters_tiny.xml", (101, 101, 1), 0.01, [1, 100], {}, []),
    "turing_uniform": ("turing_uniform.xml", (101, 101, 1), 0.07, [1, 100, 1000], {}, []),
    "turing_blob": ("turing.xml", (101, 101, 1), 0.07, [1, 10, 100], {}, []),
    "turing_50": ("turing_parameters_50x50.xml", (51, 51, 1), 0.07, [1, 100], {}, []),
    # BASELINE config 2
    "brusselator": ("The_Brusselator.xml", (101, 101, 1), 0.1, [1, 100, 1000], {}, []),
    "brusselator_300": ("The_Brusselator_300x300.xml", (301, 301, 1), 0.1, [1, 1000], {}, []),
    # BASELINE config 3
    "fish": ("fish.xml", (101, 101, 1), 1.0, [1, 10, 100], {}, []),
    "fish_300": ("fish_300x300.xml", (301, 301, 1), 1.0, [1, 20], {}, []),
    "fish_dirichlet": ("fish_big.xml", (101, 101, 1), 1.0, [1, 10, 100], {}, []),
    # BASELINE config 4
    "advection": ("advection.xml", (101, 101, 1), 0.05, [1, 10, 100, 1000], {}, ["A", "C"]),
    "transport": ("transport.xml", (101, 101, 1), 1.0, [1, 10, 100, 1000], {}, []),
    # variants that exercise paths the shipped models leave idle
    "turing_tiny_flux": ("turing_parameters_tiny.xml", (22, 22, 1), 0.07, [1, 2, 10, 100],
                         {"species_1_BC_Xp": 0.3, "species_1_BC_Ym": -0.2, "species_2_BC_Xm": 0.05, "species_2_BC_Yp": 0.11}, []),
    "brusselator_nonunit": ("The_Brusselator.xml", (64, 48, 1), 0.1, [1, 100], {}, []),  # deltaX, deltaY != 1
    "transport_rate": ("transport.xml", (101, 101, 1), 1.0, [1, 10, 100], {"membrane1": 0.5, "s0_diff": 0.25}, []),
    # synthetic models (spatial_sbml_b200/synthetic.py)
    "syn_turing_2d": ("synthetic:turing:48:40", (48, 40, 1), 0.07, [1, 10, 100], {}, []),
    "syn_turing_3d": ("synthetic:turing:20:16:12", (20, 16, 12), 0.02, [1, 10, 50], {}, []),
    "syn_brusselator_3d": ("synthetic:brusselator:14:12:10", (14, 12, 10), 0.1, [1, 20], {}, []),
    # three species: the 3- / 4-species kernel variants, stoichiometry 2, a division and a comparison in a law
    "syn_cascade3_3d": ("synthetic:cascade3:70:11:9", (70, 11, 9), 0.05, [1, 10, 40], {}, []),
    "syn_cascade3_2d": ("synthetic:cascade3:150:40", (150, 40, 1), 0.05, [1, 10, 40], {}, []),
    # 3-D boundary pass: non-zero fluxes on all six sides, Dirichlet values on the z sides
    "syn_turing_3d_flux": ("synthetic:turing_flux3d:18:14:10", (18, 14, 10), 0.02, [1, 2, 10, 40], {}, []),
    # species ON a membrane: surface diffusion over the Voronoi neighbours, binding / unbinding of a volume species, a reaction
    # between membrane species, the pseudo-membrane fill, and a membrane transport beside them (no example has a membrane species)
    "syn_turing_3d_dirichlet": ("synthetic:turing_dirichlet3d:18:14:10", (18, 14, 10), 0.02, [1, 2, 10, 40], {}, []),
    # advection with velocities that are fields (face-averaged velocity, divergence correction) / the same flow with constants
    "syn_advection_field_2d": ("synthetic:advection_field:44:36", (44, 36, 1), 0.05, [1, 2, 10, 100], {}, ["P", "Q"]),
    "syn_advection_uniform_3d": ("synthetic:advection_uniform3d:18:14:12", (18, 14, 12), 0.05, [1, 2, 10, 40], {}, ["P", "Q"]),
    "syn_advection_field_3d": ("synthetic:advection_field3d:18:14:12", (18, 14, 12), 0.05, [1, 2, 10, 40], {}, ["P", "Q"]),
    "syn_advection_uniform_2d": ("synthetic:advection_uniform:44:36", (44, 36, 1), 0.05, [1, 10, 100], {}, ["P", "Q"]),
    # assignment rules re-evaluated after every step (time-, position- and species-dependent fields, a species defined by a rule)
    "syn_rules_2d": ("synthetic:rules:30:26", (30, 26, 1), 0.05, [1, 2, 10, 100], {}, []),
    "syn_rules_3d": ("synthetic:rules:14:12:10", (14, 12, 10), 0.05, [1, 10, 40], {}, []),
    # rate rules only, with the rare operators: power, log to a base, xor, eq, neq, piecewise, tanh, cos, time, a pop-only operator
    "syn_raterules_2d": ("synthetic:raterules:30:26", (30, 26, 1), 0.05, [1, 2, 10, 100], {}, []),
    # geometry from an image (SampledFieldGeometry, y-flipped sample rows), a membrane transport across its membrane
    "syn_sampled_2d": ("synthetic:sampled:34:28", (34, 28, 1), 0.05, [1, 2, 10, 100], {}, []),
    # diffusion coefficients that are fields (the general stencil path of the stage kernel)
    "syn_fieldcoeff_2d": ("synthetic:fieldcoeff:40:30", (40, 30, 1), 0.05, [1, 10, 100], {}, []),
    "syn_fieldcoeff_3d": ("synthetic:fieldcoeff:16:12:10", (16, 12, 10), 0.05, [1, 10, 40], {}, []),
    # six staged species: the first-generation stage kernel's 8-species instantiation
    "syn_chain6_2d": ("synthetic:chain6:70:24", (70, 24, 1), 0.05, [1, 10, 100], {}, []),
    "syn_membrane_2d": ("synthetic:membrane:40:36", (40, 36, 1), 0.05, [1, 2, 10, 100], {}, []),
    "syn_membrane_mixed_2d": ("synthetic:membrane_mixed:40:36", (40, 36, 1), 0.05, [1, 2, 10, 100], {}, []),
    "syn_transport_mixed_2d": ("synthetic:transport_mixed:40:36", (40, 36, 1), 0.05, [1, 2, 10, 100], {}, []),
    "syn_membrane_box_2d": ("synthetic:membrane_box:30:26", (30, 26, 1), 0.05, [1, 2, 10, 100], {}, []),
    "syn_membrane_3d": ("synthetic:membrane:20:18:16", (20, 18, 16), 0.05, [1, 2, 10, 50], {}, []),
}


def strip_model(src_path, overrides):
    """Example model minus everything the simulator never reads."""
    for prefix, uri in (("", CORE), ("spatial", SPATIAL), ("m", MATHML)):
        ET.register_namespace(prefix, uri)
    tree = ET.parse(src_path)
    root = tree.getroot()
    drop_tags = {"{%s}annotation" % CORE, "{%s}notes" % CORE, "{%s}listOfUnitDefinitions" % CORE}
    drop_attrs = {"metaid", "name", "units", "substanceUnits", "sboTerm", "timeUnits", "volumeUnits", "areaUnits",
                  "lengthUnits", "extentUnits", "conversionFactor"}

    def clean(el):
        for child in list(el):
            if child.tag in drop_tags:
                el.remove(child)
            else:
                clean(child)
        for a in list(el.attrib):
            if a in drop_attrs:
                del el.attrib[a]
        if el.text and not el.text.strip():
            el.text = None
        if el.tail and not el.tail.strip():
            el.tail = None

    clean(root)
    for el in root.iter("{%s}parameter" % CORE):
        if el.get("id") in overrides:
            el.set("value", repr(float(overrides[el.get("id")])))
    for el in root.iter("{%s}compartment" % CORE):
        if el.get("id") in overrides:
            el.set("size", repr(float(overrides[el.get("id")])))
    text = ET.tostring(root, encoding="unicode")
    # MathML default namespace on <math> instead of a prefix on every element
    text = re.sub(r"<(/?)m:", r"<\1", text).replace('xmlns:m="%s"' % MATHML, "")
    text = text.replace("<math>", '<math xmlns="%s">' % MATHML)
    return '<?xml version="1.0" encoding="UTF-8"?>\n' + text + "\n"


def build_case(name):
    src, dims, dt, steps, overrides, advected = CASES[name]
    model_path = os.path.join(HERE, "models", name + ".xml")
    os.makedirs(os.path.dirname(model_path), exist_ok=True)
    if src.startswith("synthetic:"):
        parts = src.split(":")
        fn = getattr(synthetic, parts[1])
        text = fn(*[int(p) for p in parts[2:]])
    else:
        text = strip_model(os.path.join(EXAMPLES, src), overrides)
    with open(model_path, "w") as f:
        f.write(text)
    sbd = "/tmp/golden_%s.sbd" % name
    cmd = [DRIVER, "dump", model_path, str(dims[0]), str(dims[1]), str(dims[2]), repr(dt), sbd] + [str(s) for s in steps]
    subprocess.run(cmd, check=True)
    d = read_sbd(sbd)
    Xd, Yd, Zd, Xi, Yi, Zi = [int(v) for v in d["dims"][:6]]
    out = {"dims": np.array([Xd, Yd, Zd], dtype=np.int32), "dt": np.float64(dt), "steps": np.array(steps, dtype=np.int32),
           "delta": d["delta"], "geo_names": np.array(d["geoNames"].split()), "dimension": np.int32(d["dims"][6])}

    def even(a):
        return np.ascontiguousarray(a.reshape(Zi, Yi, Xi)[::2, ::2, ::2])

    for g in d["geoNames"].split():
        p = "geo/%s/" % g
        is_vol = int(d[p + "isVol"][0])
        out["geo/%s/isVol" % g] = np.int8(is_vol)
        dom = d[p + "isDomain"]
        bt = d[p + "bType"].reshape(-1, 6)
        if is_vol:
            out["geo/%s/isDomain" % g] = even(dom).astype(np.int8)
            out["geo/%s/isBoundary" % g] = even(d[p + "isBoundary"]).astype(np.int8)
            bits = sum((bt[:, k].astype(np.uint8) << (k + 1)) for k in range(6))
            out["geo/%s/bType" % g] = even(bits).astype(np.uint8)
            assert not dom.reshape(Zi, Yi, Xi)[1::2].any() and not dom.reshape(Zi, Yi, Xi)[:, 1::2].any() and \
                not dom.reshape(Zi, Yi, Xi)[:, :, 1::2].any(), "volume mask set off the volume cells"
        else:
            nz = np.flatnonzero(dom)
            out["geo/%s/points" % g] = nz.astype(np.int64)       # every classified point
            out["geo/%s/class" % g] = dom[nz].astype(np.int8)    # 1 membrane, 2 pseudo membrane
            out["geo/%s/bTypeAtPoints" % g] = sum((bt[nz, k].astype(np.uint8) << (k + 1)) for k in range(6)).astype(np.uint8)
            out["geo/%s/domainIndex" % g] = d[p + "domainIndex"].astype(np.int64)
            out["geo/%s/pseudoMemIndex" % g] = d[p + "pseudoMemIndex"].astype(np.int64)
            if p + "nuVec" in d:
                out["geo/%s/nuVec" % g] = d[p + "nuVec"].reshape(-1, 3)
            if p + "vorAdj" in d:  # per domainIndex entry: neighbours [XY0 YZ0 XZ0 XY1 YZ1 XZ1], s_ij, d_ij in the same slot order
                out["geo/%s/vorAdj" % g] = d[p + "vorAdj"].reshape(-1, 6).astype(np.int64)
                out["geo/%s/vorSiDi" % g] = d[p + "vorSiDi"].reshape(-1, 12)
        if p + "rpn" in d:
            out["geo/%s/rpn" % g] = np.array(d[p + "rpn"])
    i = 0
    while "rxn/%d/rpn" % i in d:
        out["rxn/%d/rpn" % i] = np.array(d["rxn/%d/rpn" % i])
        out["rxn/%d/refs" % i] = np.array(d["rxn/%d/refs" % i])
        out["rxn/%d/flags" % i] = d["rxn/%d/flags" % i]
        i += 1
    species = []
    for key in d:
        m = re.match(r"step/0/(.+)$", key)
        if m and d[key].size > 1:
            species.append(m.group(1))
    out["species"] = np.array(species)
    # a species that lives on a membrane is stored at every classified point of its membrane ("geo/<g>/points")
    on_membrane = {}
    for sp in species:
        gname = d.get("var/%s/geo" % sp, "")
        if gname and not int(d["geo/%s/isVol" % gname][0]):
            on_membrane[sp] = np.flatnonzero(d["geo/%s/isDomain" % gname])
            out["memgeo/%s" % sp] = np.array(gname)
    for s in [0] + steps:
        for sp in species:
            full = d["step/%d/%s" % (s, sp)]
            if sp in on_membrane:
                out["step/%d/%s" % (s, sp)] = full[on_membrane[sp]]
                assert not np.delete(full, on_membrane[sp]).any(), "membrane species set off its membrane"
                continue
            out["step/%d/%s" % (s, sp)] = even(full)
            if sp in advected:
                v = full.reshape(Zi, Yi, Xi)
                out["step/%d/%s/faceX" % (s, sp)] = np.ascontiguousarray(v[::2, ::2, 1::2])
                out["step/%d/%s/faceY" % (s, sp)] = np.ascontiguousarray(v[::2, 1::2, ::2])
                if Zi > 1:
                    out["step/%d/%s/faceZ" % (s, sp)] = np.ascontiguousarray(v[1::2, ::2, ::2])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    os.remove(sbd)
    print("%-22s dims %s species %s -> %.0f kB" % (name, dims, species, os.path.getsize(os.path.join(HERE, name + ".npz")) / 1e3))


if __name__ == "__main__":
    if not os.path.exists(DRIVER):
        subprocess.run(["make", "-C", os.path.join(REPO, "oracle")], check=True)
    for case in (sys.argv[1:] or list(CASES)):
        build_case(case)
