"""Probe for run-to-run differences of oracle/_ref/ref_bridge (environment size, stderr target); prints the result line of each run."""
import os
import subprocess
exe = "/root/repo/oracle/_ref/ref_bridge"
path = "/root/repo/tests/golden/models/syn_rules_2d.xml"
base = dict(os.environ)
base.pop("REF_BRIDGE_VERBOSE", None)
for label, env in [("plain", base), ("bigenv", dict(base, PYTEST_CURRENT_TEST="tests/test_parity_gpu.py::test_reference_host_setup_on_the_device_abi[syn_rules_2d] (call)")),
                   ("hugeenv", dict(base, PAD="x" * 5000))]:
    for rep in range(2):
        r = subprocess.run([exe, path, "30", "26", "1", "0.05", "10"], capture_output=True, text=True, env=env, timeout=600)
        print(label, rep, r.returncode, r.stdout.strip()[-150:])
