"""Generate tests/golden/oracle_*.npz with the CPU oracle (oracle/saeno_oracle.c, OpenMP).

These are ORACLE outputs (not reference outputs -- saenopy cannot run here):
they let the `-m gpu` tests check the CUDA path at config-1 size without paying minutes of CPU
per test run, and they let the CPU tests compare the oracle's physics with the anchors derived
from the reference's shipped lookup tables (tests/golden/reference_tables.npz).

    python tests/golden/make_oracle_golden.py          # ~10 min on 8 cores
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import c_oracle  # noqa: E402
from oracle.saeno_oracle import build_beams, lame_shell_displacement  # noqa: E402
from jointforces_b200.mesh import generate_spherical_inclusion  # noqa: E402
from jointforces_b200.simulation import spherical_boundary_conditions, radial_median_curve, lookup_bins  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
R_IN, R_OUT = 100e-6, 10000e-6
CASES = {
    # name: (lf_iso, material, pressure, step)
    "cfg1_linear_100Pa": (0.2, (2364.0, 1e30, 1e30, 1e30), 100.0, 0.0033),       # BASELINE configs[0]
    "batchA_0p1Pa": (0.2, (1449.0, 0.00215, 0.032, 0.055), 0.1, 0.0033),          # docs/data/README.md:20-42
    "linear2364_0p1Pa_step066": (0.2, (2364.0, 1e30, 1e30, 1e30), 0.1, 0.066),    # linear_reference_k2364 anchor
    "collagen12_100Pa": (0.2, (1645.0, 0.0008, 0.0075, 0.033), 100.0, 0.0033),    # BASELINE configs[1] material/load
}

if __name__ == "__main__":
    beams = build_beams(300)
    x, xc = lookup_bins()
    for name, (lf_iso, mat, p, step) in CASES.items():
        out = os.path.join(HERE, "oracle_%s.npz" % name)
        if os.path.exists(out) and "--force" not in sys.argv:
            continue
        coords, tets = generate_spherical_inclusion(lf_iso=lf_iso)
        bc = spherical_boundary_conditions(coords, R_IN, R_OUT, p)
        movable = np.any(np.isnan(bc["displacement"]), axis=1)
        t0 = time.time()
        r = c_oracle.relax(coords, tets.astype(np.int32), movable, np.nan_to_num(bc["forces"]), np.zeros_like(coords),
                           beams, mat, step, 600, 0.01, nthreads=0)
        dt = time.time() - t0
        curve = radial_median_curve(coords, r["U"], R_IN, x)
        d = np.linalg.norm(coords, axis=1)
        sel = (d > 2 * R_IN) & (d < 20 * R_IN)
        frac = float(np.median(np.linalg.norm(r["U"], axis=1)[sel] / lame_shell_displacement(d[sel], p, mat[0], R_IN, R_OUT)))
        np.savez_compressed(out, U=r["U"].astype(np.float64), relrec=r["relrec"], cg_iters=r["cg_iters"], curve=curve,
                            x_center=xc, lf_iso=lf_iso, material=np.array(mat), pressure=p, step=step,
                            n_nodes=len(coords), n_tets=len(tets), lame_fraction=frac, seconds=dt)
        print(name, "nodes", len(coords), "tets", len(tets), "newton", r["n_iter"], "cg", int(r["cg_iters"].sum()),
              "lame fraction %.4f" % frac, "%.0fs" % dt, flush=True)
