"""Developer timing of one solve (not the contract bench): python tools/quick_bench.py LF PRESSURE [MAXITER]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from jointforces_b200 import _lib, materials
from jointforces_b200.mesh import generate_spherical_inclusion
from jointforces_b200.simulation import spherical_boundary_conditions
from jointforces_b200.solver import build_beams

lf = float(sys.argv[1]) if len(sys.argv) > 1 else 0.05
pressure = float(sys.argv[2]) if len(sys.argv) > 2 else 100.0
max_iter = int(sys.argv[3]) if len(sys.argv) > 3 else 600
nsys = int(sys.argv[4]) if len(sys.argv) > 4 else 1
matname = sys.argv[5] if len(sys.argv) > 5 else "collagen12"
m = getattr(materials, matname)
mat = (m['K_0'], m['D_0'], m['L_S'], m['D_S'])
t0 = time.time(); coords, tets = generate_spherical_inclusion(length_factor=lf); t_mesh = time.time() - t0
bc = spherical_boundary_conditions(coords, 100e-6, 10000e-6, pressure)
movable = np.any(np.isnan(bc["displacement"]), axis=1)
ft = np.nan_to_num(bc["forces"])
t0 = time.time()
ctx = _lib.Context(coords, tets.astype(np.int32), movable, n_sys=nsys, beams=build_beams(300))
t_ctx = time.time() - t0
ctx.set_material(*mat)
for s in range(nsys):
    ctx.set_targets(ft * (1.0 + 0.1 * s), s); ctx.set_displacements(np.zeros_like(coords), s)
info = ctx.info()
print("mesh %.1fs ctx %.2fs" % (t_mesh, t_ctx), info)
for rep in range(2):
    for s in range(nsys):
        ctx.set_displacements(np.zeros_like(coords), s)
    ctx.reset_stats()
    t0 = time.time(); relrec, cgit, n = ctx.relax(0.0033, max_iter, 0.01); wall = time.time() - t0
    st = ctx.stats()
    print("rep", rep, "wall %.3fs" % wall, "newton", n.tolist(), "cg", [int(c.sum()) for c in cgit],
          "ms el %.2f as %.2f cg %.2f" % (st['ms_element'], st['ms_assembly'], st['ms_cg']))
    print("  tet-updates/s %.3e  cg-iters/s %.3e  us/cg-iter %.2f  us/tet %.4f" % (
        st['tet_updates'] / (st['ms_element'] * 1e-3), st['cg_iterations'] / (st['ms_cg'] * 1e-3),
        st['ms_cg'] * 1e3 / max(1, max(int(c.sum()) for c in cgit)), st['ms_element'] * 1e3 / st['tet_updates']))
    # algorithmic bytes per CG iteration: matrix (76 B per block incl. column index) + 8 vector passes of 32 B per row
    by = info['n_blocks'] * 76 + info['n_free'] * (8 * 32)
    print("  cg algorithmic GB/s %.1f" % (by * st['cg_iterations'] / (st['ms_cg'] * 1e-3) / 1e9))
print("fp64 peak GF", _lib.bench_fp64())
