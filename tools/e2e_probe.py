"""Where does the host-facing call spend its time?  JF_TIMING=1 python tools/e2e_probe.py [cfg]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import build_case, SOLVER
from jointforces_b200 import _lib
case = build_case(sys.argv[1] if len(sys.argv) > 1 else "cfg2")
_lib.device_count()
for rep in range(3):
    t0 = time.time()
    ctx = _lib.Context(case["coords"], case["tets"], case["movable"], n_sys=1, beams=case["beams"]); t1 = time.time()
    ctx.set_material(*case["mat"]); ctx.set_targets(case["ft"]); ctx.set_displacements(np.zeros_like(case["coords"])); t2 = time.time()
    relrec, cgit, n = ctx.relax(SOLVER["step"], SOLVER["max_iter"], SOLVER["conv_crit"]); t3 = time.time()
    U = ctx.displacements(); f = ctx.forces(); t4 = time.time()
    ctx.close(); t5 = time.time()
    print("rep %d: create %.1f ms, set %.1f ms, relax %.1f ms, get %.1f ms, destroy %.1f ms" % (rep, *(1e3 * np.diff([t0, t1, t2, t3, t4, t5]))), flush=True)
