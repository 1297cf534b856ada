"""Short streaming-CG run for ncu on an HBM-bound matrix: python tools/profile_cg_stream.py [cfg5] [maxiter]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import build_case, cg_bytes_per_iteration
from jointforces_b200 import _lib
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
maxiter = int(sys.argv[2]) if len(sys.argv) > 2 else 50
case = build_case(wl)
ctx = _lib.Context(case["coords"], case["tets"], case["movable"], n_sys=1, beams=case["beams"])
ctx.set_material(*case["mat"]); ctx.set_targets(case["ft"]); ctx.set_displacements(np.zeros_like(case["coords"]))
ctx.evaluate()
ctx.cg(1e-30, maxiter)            # warm-up launch
ctx.reset_stats()
it = ctx.cg(1e-30, maxiter)
st, info = ctx.stats(), ctx.info()
by = cg_bytes_per_iteration(info)
print(wl, info)
print("cg iters", it.tolist(), "ms_cg %.3f  us/iter %.2f  algorithmic GB/s %.1f  bytes/iter %.4g" % (
    st["ms_cg"], st["ms_cg"] * 1e3 / it[0], float(by) * float(it[0]) / (st["ms_cg"] * 1e-3) / 1e9, by))
