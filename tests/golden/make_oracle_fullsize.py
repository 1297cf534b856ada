"""Generate tests/golden/oracle_full_*.npz: CPU-oracle solutions on the BASELINE meshes at FULL size.

ORACLE outputs (oracle/saeno_oracle.c, OpenMP), not reference outputs -- saenopy cannot run here
(SURVEY.md §0.2), parity with real saenopy stays unpinned.  They let the `-m gpu` tests compare
every benchmarked kernel path with the oracle on the meshes the benchmark numbers are quoted on
(tests/test_gpu_fullsize.py) without paying tens of minutes of CPU per test run:

  cfg2_full        BASELINE configs[1]: lf 0.05 (17,612 nodes / 102,744 tets), collagen12, 100 Pa, whole relaxation
  cfg3_row{A,B}    two more rows of the configs[2] sweep on the same mesh (0.1 Pa and 10 kPa), whole relaxation
  cfg4_full        configs[3]: lf 0.03 (78,057 / 460,172), collagen24, 10 kPa, whole relaxation
  cfg5_loose       configs[4]: lf 0.015 (603,402 / 3,588,654), collagen06, 100 Pa, first 3 Newton steps, saenopy's cg tol
  cfg5_tight       same, first 2 Newton steps, cg tol 1e-10 (the stop iteration no longer matters)

To keep the files small the displacement field of the larger meshes is stored on a fixed strided
subset of the nodes (`stride`; node i is kept when i % stride == 0); curve, relrec and CG iteration
counts are stored whole.

    python tests/golden/make_oracle_fullsize.py [name ...]     # ~1.5 h on 6 threads for all
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import c_oracle  # noqa: E402
from oracle.saeno_oracle import build_beams  # noqa: E402
from jointforces_b200.mesh import generate_spherical_inclusion  # noqa: E402
from jointforces_b200.simulation import spherical_boundary_conditions, radial_median_curve, lookup_bins  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
R_IN, R_OUT = 100e-6, 10000e-6
C06, C12, C24 = (447.0, 0.0008, 0.0075, 0.033), (1645.0, 0.0008, 0.0075, 0.033), (5208.0, 0.0008, 0.0075, 0.033)
CASES = {
    # name: (length_factor, material, pressure, max_iter, cg_tol (0 = saenopy's 1e-5), stride of the stored U)
    "cfg2_full": (0.05, C12, 100.0, 600, 0.0, 1),
    "cfg3_rowA": (0.05, C12, 0.1, 600, 0.0, 4),
    "cfg3_rowB": (0.05, C12, 10000.0, 600, 0.0, 4),
    "cfg4_full": (0.03, C24, 10000.0, 600, 0.0, 8),
    "cfg5_loose": (0.015, C06, 100.0, 3, 0.0, 64),
    "cfg5_tight": (0.015, C06, 100.0, 2, 1e-10, 64),
}

if __name__ == "__main__":
    beams = build_beams(300)
    x, xc = lookup_bins()
    names = [a for a in sys.argv[1:] if not a.startswith("-")] or list(CASES)
    threads = int(os.environ.get("ORACLE_THREADS", "6"))
    for name in names:
        lf, mat, p, max_iter, cg_tol, stride = CASES[name]
        out = os.path.join(HERE, "oracle_full_%s.npz" % name)
        if os.path.exists(out) and "--force" not in sys.argv:
            continue
        coords, tets = generate_spherical_inclusion(length_factor=lf)
        bc = spherical_boundary_conditions(coords, R_IN, R_OUT, p)
        movable = np.any(np.isnan(bc["displacement"]), axis=1)
        t0 = time.time()
        r = c_oracle.relax(coords, tets.astype(np.int32), movable, np.nan_to_num(bc["forces"]), np.zeros_like(coords),
                           beams, mat, 0.0033, max_iter, 0.01, nthreads=threads, cg_tol=cg_tol)
        dt = time.time() - t0
        curve = radial_median_curve(coords, r["U"], R_IN, x)
        np.savez_compressed(out, U_sub=r["U"][::stride].astype(np.float64), stride=stride, U_norm=np.linalg.norm(r["U"]),
                            relrec=r["relrec"], cg_iters=r["cg_iters"], curve=curve, x_center=xc, length_factor=lf,
                            material=np.array(mat), pressure=p, step=0.0033, max_iter=max_iter, cg_tol=cg_tol,
                            n_nodes=len(coords), n_tets=len(tets), seconds=dt, threads=threads)
        print(name, "nodes", len(coords), "tets", len(tets), "newton", r["n_iter"], "cg", int(r["cg_iters"].sum()),
              "%.0fs" % dt, "element %.0fs cg %.0fs" % (r["t_element"], r["t_cg"]), flush=True)
