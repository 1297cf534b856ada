"""Short run for ncu: a few Newton iterations of one workload.  python tools/profile_run.py cfg2 3"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import build_case, SOLVER
from jointforces_b200 import _lib
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
case = build_case(wl)
ctx = _lib.Context(case["coords"], case["tets"], case["movable"], n_sys=1, beams=case["beams"])
ctx.set_material(*case["mat"]); ctx.set_targets(case["ft"]); ctx.set_displacements(np.zeros_like(case["coords"]))
relrec, cgit, n_iter = ctx.relax(SOLVER["step"], n, SOLVER["conv_crit"])
st = ctx.stats()
print(wl, ctx.info())
print("newton", n_iter, "cg", cgit[0].tolist(), {k: round(v, 3) if isinstance(v, float) else v for k, v in st.items()})
