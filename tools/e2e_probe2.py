import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import build_case, SOLVER
from jointforces_b200 import _lib
from jointforces_b200.simulation import radial_median_curve, lookup_bins
case = build_case("cfg2"); x,_ = lookup_bins()
for rep in range(4):
    t0=time.time()
    out=_lib.solve_boundarycondition(case["coords"], case["tets"], case["movable"], case["ft"], np.zeros_like(case["coords"]), case["mat"], SOLVER["step"], SOLVER["max_iter"], SOLVER["conv_crit"], beams=case["beams"])
    t1=time.time(); c=radial_median_curve(case["coords"], out["U"], case["r_in"], x); t2=time.time()
    st=out["stats"]
    print("rep %d total %.1f ms: call %.1f (setup %.1f relax %.1f el %.1f as %.1f cg %.1f) curve %.1f"%(rep,(t2-t0)*1e3,(t1-t0)*1e3,st["ms_setup"],st["ms_relax"],st["ms_element"],st["ms_assembly"],st["ms_cg"],(t2-t1)*1e3), flush=True)
