"""Does a working warm start explain the high-pressure rows of the reference's shipped table?
Runs the BatchA recipe (docs/data/README.md:20-42) on the 10,261-node synthetic mesh, cold and warm, and
prints median(ours / shipped) per row.  Needs tests/golden/reference_tables.npz.  One GPU."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import jointforces_b200 as jf
from jointforces_b200.sweep import relax_load_cases
from jointforces_b200.mesh import generate_spherical_inclusion
ref = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "reference_tables.npz"))
rows = list(ref["rows"]); x = ref["x"]
coords, tets = generate_spherical_inclusion(lf_iso=0.2)
r_in = 100 * 10 ** -6
pressures = np.logspace(-1, 4, 150)
bcs = [jf.simulation.spherical_boundary_conditions(coords, r_in, 10000 * 10 ** -6, float(p)) for p in pressures]
edges, _ = jf.simulation.lookup_bins()
mat = jf.materials.custom(1449, 0.00215, 0.032, 0.055)
for warm in (False, True):
    t0 = time.time()
    res, st = relax_load_cases(coords, tets, bcs[0]["displacement"], [b["forces"] for b in bcs], mat, 0.0033, 600, 0.01,
                               curve_bins=(r_in, edges), fetch_fields=False, warm_start=warm, batch=1)
    print("warm_start=%s: %.1f s, Newton steps %d, CG iterations %d" % (warm, time.time() - t0, st["newton_steps"], st["cg_iterations"]))
    for row in rows:
        y = res[row]["curve"]; ok = ~np.isnan(y) & (x > 1.2) & (x < 40)
        near = ok & (x < 3); far = ok & (x > 10)
        print("  row %3d p=%9.3g newton %3d  ours/shipped: all %.3f near %.3f far %.3f" % (
            row, pressures[row], len(res[row]["relrec"]) - 1, np.median(y[ok] / ref["batchA_y"][rows.index(row)][ok]),
            np.median(y[near] / ref["batchA_y"][rows.index(row)][near]), np.median(y[far] / ref["batchA_y"][rows.index(row)][far])))
