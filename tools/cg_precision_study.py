"""Why do two correct implementations of saenopy's CG stop at different iterations?  

The CUDA kernels take 0.6-2.8 % FEWER CG iterations per Newton step than oracle/saeno_oracle.c on the BASELINE
config-2 mesh, for the classic two-reduction kernels exactly as for the single-reduction ones (tools/cg_iter_bias.py),
so the recurrence for r.r is not the cause.  This script runs the oracle's own CG on the oracle's own matrix three
times, changing nothing but how sums are evaluated: plain double (the oracle), FMA + pairwise sums (what a GPU does),
long double.  CPU only:  python tools/cg_precision_study.py [pressure] [newton_steps_before]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import c_oracle
from oracle.saeno_oracle import build_beams
from jointforces_b200.mesh import generate_spherical_inclusion
from jointforces_b200.simulation import spherical_boundary_conditions

p = float(sys.argv[1]) if len(sys.argv) > 1 else 10000.0
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
mat = (1645.0, 0.0008, 0.0075, 0.033)
coords, tets = generate_spherical_inclusion(length_factor=0.05)
bc = spherical_boundary_conditions(coords, 100e-6, 10000e-6, p)
movable = np.any(np.isnan(bc["displacement"]), axis=1)
ft = np.nan_to_num(bc["forces"])
beams = build_beams(300)
r = c_oracle.relax(coords, tets.astype(np.int32), movable, ft, np.zeros_like(coords), beams, mat, 0.0033, steps, 0.01,
                   nthreads=int(os.environ.get("ORACLE_THREADS", "2")))
print("pressure %g Pa, state after %d Newton steps (oracle CG iterations so far: %s)" % (p, steps, r["cg_iters"].tolist()))
print("next CG solve, iterations by arithmetic:", c_oracle.cg_variants(coords, tets.astype(np.int32), movable, ft, r["U"], beams, mat))
