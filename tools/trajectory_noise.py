import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from conftest import make_case, COLLAGEN12
from jointforces_b200 import _lib
from jointforces_b200.solver import build_beams
from oracle import c_oracle
beams = build_beams(300)
for p in (1.0, 100.0, 3000.0):
    c = make_case(0.667, p)
    def run():
        return _lib.solve_boundarycondition(c["coords"], c["tets"], c["movable"], c["f_target"], np.zeros_like(c["coords"]), COLLAGEN12, 0.0033, 600, 0.01, beams=beams)
    a = run()
    os.environ["JF_CG_SLOT_SUMS"] = "1"; b = run(); del os.environ["JF_CG_SLOT_SUMS"]
    os.environ["JF_CG_CLASSIC"] = "1"; d = run(); del os.environ["JF_CG_CLASSIC"]
    o = c_oracle.relax(c["coords"], c["tets"], c["movable"], c["f_target"], np.zeros_like(c["coords"]), beams, COLLAGEN12, 0.0033, 600, 0.01, nthreads=1)
    o8 = c_oracle.relax(c["coords"], c["tets"], c["movable"], c["f_target"], np.zeros_like(c["coords"]), beams, COLLAGEN12, 0.0033, 600, 0.01, nthreads=8)
    e = lambda x, y: np.linalg.norm(x["U"] - y["U"]) / np.linalg.norm(y["U"])
    print("p=%g newton fx %d slot %d classic %d oracle %d | cg fx %d slot %d classic %d oracle %d oracle8 %d" % (p, a["n_iter"], b["n_iter"], d["n_iter"], o["n_iter"],
          a["cg_iters"].sum(), b["cg_iters"].sum(), d["cg_iters"].sum(), o["cg_iters"].sum(), o8["cg_iters"].sum()))
    print("   err fx-slot %.2e fx-classic %.2e slot-classic %.2e | fx-oracle %.2e slot-oracle %.2e classic-oracle %.2e oracle8-oracle1 %.2e" % (
        e(a, b), e(a, d), e(b, d), e(a, o), e(b, o), e(d, o), e(o8, o)))
