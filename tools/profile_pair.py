"""Two load cases side by side on the configs[1] mesh for a few Newton steps (ncu target for jf_cg_pair_kernel):
    python tools/profile_pair.py [newton_steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from jointforces_b200 import _lib, materials
from jointforces_b200.mesh import generate_spherical_inclusion
from jointforces_b200.simulation import spherical_boundary_conditions
from jointforces_b200.solver import build_beams

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
m = materials.collagen12
coords, tets = generate_spherical_inclusion(length_factor=0.05)
bc = spherical_boundary_conditions(coords, 100e-6, 10000e-6, 100.0)
movable = np.any(np.isnan(bc["displacement"]), axis=1)
ft = np.nan_to_num(bc["forces"])
with _lib.Context(coords, tets.astype(np.int32), movable, n_sys=2, beams=build_beams(300)) as ctx:
    ctx.set_material(m['K_0'], m['D_0'], m['L_S'], m['D_S'])
    for s, f in enumerate((1.0, 1.1)):
        ctx.set_targets(ft * f, s)
        ctx.set_displacements(np.zeros_like(coords), s)
    relrec, cgit, n = ctx.relax(0.0033, steps, 0.01)
    print("pair", ctx.info())
    print("newton", n.tolist(), "cg", [c.tolist() for c in cgit], ctx.stats())
