"""A/B of the CG kernels on an HBM-bound matrix (503 k nodes icoshell): python tools/hbm_cg_ab.py TOL [MAXITER]
tol >= 1e-9 lets the single-reduction kernels run (if configured, e.g. JF_CG_FORCE_STREAM_HALO=1), tol below selects the classic ones."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import cg_bytes_per_iteration
from jointforces_b200 import _lib, materials
from jointforces_b200.mesh import generate_spherical_inclusion
from jointforces_b200.simulation import spherical_boundary_conditions
from jointforces_b200.solver import build_beams
tol = float(sys.argv[1]) if len(sys.argv) > 1 else 1e-30
maxiter = int(sys.argv[2]) if len(sys.argv) > 2 else 50
coords, tets = generate_spherical_inclusion(lf_iso=0.05, method="icoshell")
bc = spherical_boundary_conditions(coords, 100 * 10 ** -6, 10000 * 10 ** -6, 100.0)
movable = np.any(np.isnan(bc["displacement"]), axis=1)
m = materials.collagen06
for env in ({}, {"JF_CG_FORCE_STREAM_HALO": "1"}):
    for k in ("JF_CG_FORCE_STREAM_HALO",):
        os.environ.pop(k, None)
    os.environ.update(env)
    with _lib.Context(coords, tets.astype(np.int32), movable, n_sys=1, beams=build_beams(300)) as ctx:
        ctx.set_material(m['K_0'], m['D_0'], m['L_S'], m['D_S']); ctx.set_targets(np.nan_to_num(bc["forces"]))
        ctx.set_displacements(np.zeros_like(coords)); ctx.evaluate()
        info = ctx.info()
        for t in (1e-30, tol):
            ctx.cg(t, maxiter); ctx.reset_stats()
            it = [int(ctx.cg(t, maxiter)[0]) for _ in range(3)]
            st = ctx.stats()
            by = cg_bytes_per_iteration(info)
            print(env, "tol %g" % t, "halo %d fused %d" % (info["cg_stream_halo"], info["cg_fused"]), "iters", it,
                  "us/iter %.2f  GB/s %.0f" % (st["ms_cg"] * 1e3 / sum(it), by * sum(it) / (st["ms_cg"] * 1e-3) / 1e9), flush=True)
