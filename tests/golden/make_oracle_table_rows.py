"""Oracle curves for the reference's BatchA recipe (docs/data/README.md:20-42: K_0=1449, D_0=0.00215, L_S=0.032,
D_S=0.055, step 0.0033, max_iter 600) at the table rows kept in reference_tables.npz, on the 10,261-node synthetic
mesh.  ORACLE outputs (saenopy cannot run here); compared in tests/test_oracle.py with the reference's shipped
Collagen12_BatchA.npy rows (different mesh => few-% agreement) and in the -m gpu tests with the CUDA path (1e-3).

    python tests/golden/make_oracle_table_rows.py        # ~20 min on 8 cores
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
MAT = (1449.0, 0.00215, 0.032, 0.055)
ROWS = [30, 60, 90, 120, 149]

if __name__ == "__main__":
    out = os.path.join(HERE, "oracle_batchA_rows.npz")
    pressures = np.logspace(-1, 4, 150)
    coords, tets = generate_spherical_inclusion(lf_iso=0.2)
    beams = build_beams(300)
    x, xc = lookup_bins()
    r_in = 100 * 10 ** -6
    curves, n_iters, cg_totals, energies = [], [], [], []
    for row in ROWS:
        p = float(pressures[row])
        bc = spherical_boundary_conditions(coords, r_in, 10000 * 10 ** -6, p)
        movable = np.any(np.isnan(bc["displacement"]), axis=1)
        t0 = time.time()
        r = c_oracle.relax(coords, tets.astype(np.int32), movable, np.nan_to_num(bc["forces"]), np.zeros_like(coords),
                           beams, MAT, 0.0033, 600, 0.01, nthreads=0)
        curves.append(radial_median_curve(coords, r["U"], r_in, x))
        n_iters.append(r["n_iter"]); cg_totals.append(int(r["cg_iters"].sum())); energies.append(r["relrec"][-1, 1])
        print("row", row, "p=%.4g" % p, "newton", r["n_iter"], "cg", cg_totals[-1], "%.0fs" % (time.time() - t0), flush=True)
        np.savez_compressed(out, rows=np.array(ROWS[:len(curves)]), pressure=pressures[ROWS[:len(curves)]], curves=np.array(curves),
                            n_iter=np.array(n_iters), cg_total=np.array(cg_totals), energy=np.array(energies), x_center=xc,
                            material=np.array(MAT), lf_iso=0.2, step=0.0033)
