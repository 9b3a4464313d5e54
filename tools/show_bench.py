import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d.get("roofline") or {}
print("ms/step %.3f  value %.3e  frac %.3f  achieved %.0f GB/s  e2e %.3e  launches %s  clk %s" % (
    d["ms_per_step"], d["value"], r.get("frac", 0), r.get("achieved", 0), (d.get("e2e") or {}).get("value", 0), d.get("gpu_launches"), d.get("clocks")))
