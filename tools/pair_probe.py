"""Timing probe of jf_cg_pair_kernel on the configs[1] mesh: both systems busy / only one busy (the other system's right-hand
side is zero, so its CTAs leave their CG loop at once and the busy system has its SMs to itself) / the single-solve kernel.
    python tools/pair_probe.py [newton_steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from jointforces_b200 import _lib, materials
from jointforces_b200.mesh import generate_spherical_inclusion
from jointforces_b200.simulation import spherical_boundary_conditions
from jointforces_b200.solver import build_beams

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
m = materials.collagen12
mat = (m['K_0'], m['D_0'], m['L_S'], m['D_S'])
coords, tets = generate_spherical_inclusion(length_factor=0.05)
bc = spherical_boundary_conditions(coords, 100e-6, 10000e-6, 100.0)
movable = np.any(np.isnan(bc["displacement"]), axis=1)
ft = np.nan_to_num(bc["forces"])
for label, nsys, f1 in (("single-solve kernel", 1, None), ("pair, both busy", 2, 1.1), ("pair, one busy", 2, 0.0)):
    with _lib.Context(coords, tets.astype(np.int32), movable, n_sys=nsys, beams=build_beams(300)) as ctx:
        ctx.set_material(*mat)
        for rep in range(2):
            ctx.set_targets(ft, 0); ctx.set_displacements(np.zeros_like(coords), 0)
            if nsys == 2:
                ctx.set_targets(ft * f1, 1); ctx.set_displacements(np.zeros_like(coords), 1)
            ctx.reset_stats()
            relrec, cgit, n = ctx.relax(0.0033, steps, 0.01)
            st = ctx.stats()
        its = [int(c.sum()) for c in cgit]
        print("%-20s cg_fused %d  newton %s  cg iterations %s  ms_cg %.2f  us per iteration of the longest system %.3f, per system-iteration %.3f"
              % (label, ctx.info()["cg_fused"], n.tolist(), its, st["ms_cg"], st["ms_cg"] * 1e3 / max(its), st["ms_cg"] * 1e3 / max(1, sum(its))))
