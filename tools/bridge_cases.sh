cd $GRAFT_REPO_ROOT
python - <<'PY'
import numpy as np, subprocess, json
cases=["turing_tiny","brusselator","turing_blob","fish","fish_dirichlet","turing_tiny_flux","syn_turing_3d","syn_turing_3d_flux","syn_raterules_2d","syn_fieldcoeff_2d","syn_fieldcoeff_3d","syn_chain6_2d","syn_cascade3_3d","syn_sampled_2d","advection"]
for c in cases:
    g=np.load("tests/golden/%s.npz"%c); d=[int(v) for v in g["dims"]]
    r=subprocess.run(["oracle/_ref/ref_bridge","tests/golden/models/%s.xml"%c,str(d[0]),str(d[1]),str(d[2]),repr(float(g["dt"])),"10"],capture_output=True,text=True)
    print(c, r.returncode, (r.stdout.strip() or r.stderr.strip())[-230:])
PY
