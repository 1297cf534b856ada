"""Tiny end-to-end run for compute-sanitizer: all kernel families on small inputs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from jointforces_b200 import _lib
from jointforces_b200.mesh import generate_spherical_inclusion
from jointforces_b200.simulation import spherical_boundary_conditions, lookup_bins
coords, tets = generate_spherical_inclusion(lf_iso=0.667)
bc = spherical_boundary_conditions(coords, 1e-4, 1e-2, 100.0)
mov = np.any(np.isnan(bc["displacement"]), axis=1); ft = np.nan_to_num(bc["forces"])
x, _ = lookup_bins()
for flags, env in ((0, None), (0, "JF_CG_NO_REGRES"), (_lib.JF_FLAG_NO_RESIDENT, None), (_lib.JF_FLAG_NO_RESIDENT, "JF_CG_NO_STREAM_HALO")):
    if env: os.environ[env] = "1"
    with _lib.Context(coords, tets.astype(np.int32), mov, n_sys=2, flags=flags) as ctx:
        ctx.set_material(1645, 0.0008, 0.0075, 0.033)
        for s in range(2):
            ctx.set_targets(ft * (1 + s), s); ctx.set_displacements(np.zeros_like(coords), s)
        ctx.set_curve_bins(coords, 1e-4, x)
        relrec, cgit, n = ctx.relax(0.0033, 4, 0.01)
        c = ctx.curves(); y = ctx.apply_stiffness(np.ones_like(coords)); it = ctx.cg(1e-5, 10)
        print(ctx.info()["cg_resident"], ctx.info()["cg_stream_halo"], n.tolist(), [int(v.sum()) for v in cgit], float(np.nanmax(c)))
    if env: del os.environ[env]
