#!/usr/bin/env python
"""TMEM load/store throughput as the epilogue warps see it (16 warps of one CTA): python scripts/tmem_bench.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import density_planner_b200 as dp
from density_planner_b200 import _lib as _dl
L = _dl.diag_lib()
for mode, name in ((0, "tcgen05.ld x16"), (1, "tcgen05.st x16"), (2, "ld -> st pair")):
    c = ctypes.c_double(0)
    assert L.dp_tc_tmem_bench(0, mode, 2000, ctypes.byref(c)) == 0
    # 16 warps each issue one 2 KB instruction per `c` cycles
    print("%-16s %.1f cycles per warp instruction with 16 warps in flight -> %.0f B/cycle/SM" % (name, c.value, 16 * 2048 / c.value))
