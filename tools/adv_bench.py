"""Times the CIP-CSLR direction passes at n x n (sb2_time_advection).  usage: python tools/adv_bench.py [n] [reps]"""
import ctypes as C, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import spatial_sbml_b200 as sb
from spatial_sbml_b200 import synthetic
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
sim = sb.SpatialSimulator()
sim.initFromString(synthetic.advection_uniform(n, n), n, n, 1)
lib = sb.load_library()
ms, npass = C.c_double(0.0), C.c_int(0)
rc = lib.sb2_time_advection(C.c_void_p(sim.getDeviceHandle()), 0.05, reps, C.byref(ms), C.byref(npass))
cells = sim.getNumDomainCells()
print("rc", rc, "n", n, "ms/pass %.4f" % ms.value, "passes", npass.value, "GB/s(40B) %.0f" % (40.0 * cells / (ms.value * 1e-3) / 1e9))
sim.close()
