#!/usr/bin/env python
"""Writes tests/golden/stability.json by calling the REFERENCE's own checkDiffusionStab / checkAdvectionStab
(checkStability.cpp:10-52, 119-154) through oracle/_ref/ref_driver stab, for every golden model and a few step sizes.
Runs in the build container only (the binary is built from /root/reference)."""
import json
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
DRIVER = os.path.join(REPO, "oracle", "_ref", "ref_driver")

out = {}
for f in sorted(os.listdir(HERE)):
    if not f.endswith(".npz"):
        continue
    case = f[:-4]
    g = np.load(os.path.join(HERE, f))
    dims = [str(int(v)) for v in g["dims"]]
    dt = float(g["dt"])
    dts = [repr(dt), repr(dt * 10), repr(dt * 100), "1e-3"]
    r = subprocess.run([DRIVER, "stab", os.path.join(HERE, "models", case + ".xml")] + dims + dts, capture_output=True, text=True, check=True)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    out[case] = json.loads(line)
    print(case, out[case])
with open(os.path.join(HERE, "stability.json"), "w") as fh:
    json.dump(out, fh, indent=1, sort_keys=True)
