import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import density_planner_b200 as dp
from density_planner_b200 import _lib as _dl
L = _dl.diag_lib()
out = (ctypes.c_double * 2)()
for variant in (0, 1, 2, 3):
    for n in (32, 48, 64, 128, 256):
        rc = L.dp_tc_mma_bench(0, variant, n, 500, out)
        assert rc == 0, L.dp_last_error()
        print("variant %d N=%3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA  -> %.0f MAC/cyc" % (variant, n, out[0], out[1], 128 * n * 16 / out[1]))
