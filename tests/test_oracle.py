"""CPU tests of the oracle itself: it has to be trustworthy before it may judge the CUDA path.
Pins: finite differences of its own energy, the strain.py:336-338 stiffness law, the exact Lame
shell solution, and the anchors derived from the reference's shipped tables (tests/golden)."""
import os

import numpy as np
import pytest

from conftest import COLLAGEN12, LINEAR, BATCH_A, make_case
from oracle import c_oracle
from oracle.saeno_oracle import (SemiAffineFiberMaterial, Solver, build_beams, cg, fold_beams,
                                 lame_shell_displacement)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_beams_match_saenopy_recipe():
    s = build_beams(300)
    assert s.shape == (286, 3)
    assert np.allclose(np.linalg.norm(s, axis=1), 1.0, atol=1e-15)
    m2 = (s[:, :, None] * s[:, None, :]).mean(0)
    assert np.allclose(np.diag(m2), [0.3357, 0.3357, 0.3286], atol=1e-4)
    assert np.abs(m2 - np.diag(np.diag(m2))).max() < 1e-15
    d, w = fold_beams(s)
    assert len(d) == 172 and w.sum() == 286 and set(w) == {1.0, 2.0}       # two 29-point rings do not pair


def test_stiffness_law_three_regimes():
    """w'' exactly as the commented formula at reference strain.py:336-338."""
    k, d0, ls, ds = COLLAGEN12
    mat = SemiAffineFiberMaterial(k, d0, ls, ds)
    lam = np.array([-0.01, -1e-4, 0.0, 0.003, ls, 0.02, 0.1])
    expect = np.where(lam < 0, k * np.exp(lam / d0), np.where(lam < ls, k, k * np.exp((lam - ls) / ds)))
    assert np.allclose(mat(lam)[2], expect, rtol=1e-14)
    assert np.allclose(c_oracle.material_eval(k, d0, ls, ds, lam)[2], expect, rtol=1e-14)


@pytest.mark.parametrize("mat", [COLLAGEN12, BATCH_A, (2091.0, 0.002, 1e30, 1e30)])
def test_material_integrals(mat):
    """w' and w are the integrals of w'' with w(0) = w'(0) = 0 and are continuous at the regime joints."""
    m = SemiAffineFiberMaterial(*mat)
    lam = np.linspace(-0.02, 0.08, 200001)
    w0, w1, w2 = m(lam)
    i0 = np.argmin(np.abs(lam))
    num1 = np.concatenate([[0], np.cumsum(0.5 * (w2[1:] + w2[:-1]) * np.diff(lam))])
    num1 -= num1[i0]
    num0 = np.concatenate([[0], np.cumsum(0.5 * (w1[1:] + w1[:-1]) * np.diff(lam))])
    num0 -= num0[i0]
    assert np.abs(w1 - num1).max() < 1e-6 * np.abs(w1).max()
    assert np.abs(w0 - num0).max() < 1e-6 * np.abs(w0).max()
    assert m(np.array([0.0]))[0][0] == 0 and m(np.array([0.0]))[1][0] == 0


def test_linear_material_has_no_cancellation():
    """materials.linear sets d0 = ls = ds = 1e30 (reference materials.py:8): w must be k lam^2/2 to rounding."""
    m = SemiAffineFiberMaterial(*LINEAR)
    lam = np.array([-0.3, -1e-3, -1e-9, 1e-9, 1e-3, 0.3])
    w0, w1, w2 = m(lam)
    assert np.allclose(w0, 0.5 * LINEAR[0] * lam ** 2, rtol=1e-14)
    assert np.allclose(w1, LINEAR[0] * lam, rtol=1e-14)
    assert np.allclose(w2, LINEAR[0], rtol=1e-14)
    c0, c1, c2 = c_oracle.material_eval(*LINEAR, lam)
    assert np.allclose(c0, w0, rtol=1e-15) and np.allclose(c1, w1, rtol=1e-15)


def test_lookup_table_emulation_is_close_to_closed_form():
    """saenopy samples w'' at 1e-6 and integrates by cumulative sums; which index rule it uses is
    not known, so both are exercised.  Away from lambda = 0 the table agrees with the
    closed forms to about step/lambda."""
    lam = np.array([-0.05, -0.004, 0.0031, 0.0102, 0.05, 0.3])
    closed = SemiAffineFiberMaterial(*COLLAGEN12)(lam)
    for rule in ("ceil", "round", "floor", "interp"):
        lut = SemiAffineFiberMaterial(*COLLAGEN12, mode="lut", lut_index=rule)(lam)
        assert np.allclose(lut[2], closed[2], rtol=2e-3)
        assert np.allclose(lut[1], closed[1], rtol=2e-3)
        assert np.allclose(lut[0], closed[0], rtol=4e-3)


def test_lookup_table_interpolation_rule():
    """The lookup rule recalled for saenopy (default of mode='lut'): linear interpolation between the samples
    floor(q) and floor(q)+1.  It reproduces the sampled tables exactly at the samples, lies between the two
    single-sample rules in between, is continuous in lambda (the single-sample rules jump by one table step)
    and is an order of magnitude closer to the closed forms than they are for w'."""
    m = SemiAffineFiberMaterial(*COLLAGEN12, mode="lut")
    assert m.lut_index == "interp"
    fl = SemiAffineFiberMaterial(*COLLAGEN12, mode="lut", lut_index="floor")
    ce = SemiAffineFiberMaterial(*COLLAGEN12, mode="lut", lut_index="ceil")
    fl._lut = ce._lut = m._lut if m._lut is not None else (m._build_lut() or m._lut)
    on_grid = -1.0 + 1e-6 * np.array([1000123, 1010500, 1200000], dtype=np.float64)
    q = (on_grid + 1.0) / 1e-6
    on_grid = on_grid[q == np.floor(q)]
    assert len(on_grid)
    for a, b in zip(m(on_grid), fl(on_grid)):
        assert np.array_equal(a, b)
    lam = np.array([-0.0300004, 0.0040007, 0.0200003, 0.2500006])
    lo, hi, mid = fl(lam), ce(lam), m(lam)
    for k in range(3):
        assert np.all((mid[k] >= np.minimum(lo[k], hi[k])) & (mid[k] <= np.maximum(lo[k], hi[k])))
    eps = 1e-9
    for k in range(3):   # continuity across a sample
        a, b = m(np.array([0.02 - eps]))[k], m(np.array([0.02 + eps]))[k]
        assert abs(a - b) <= 1e-6 * abs(a) + 1e-18
    closed = SemiAffineFiberMaterial(*COLLAGEN12)(lam)
    # the cumulative sums are Riemann sums: w' carries an offset of about step * w''/2 whatever the rule
    assert np.allclose(mid[1], closed[1], rtol=1e-3)
    # far end of the table: clamped, no index error
    w = m(np.array([3.9999999, 4.5, -1.0]))
    assert all(np.all(np.isfinite(x)) for x in w)


def _solver(case, mat):
    S = Solver()
    S.set_material_model(SemiAffineFiberMaterial(*mat))
    S.set_nodes(case["coords"])
    S.set_tetrahedra(case["tets"].astype(float))   # the reference passes floats (simulation.py:71-76)
    S.set_boundary_condition(case["bc"]["displacement"], case["bc"]["forces"])
    return S


def test_force_is_energy_gradient_and_K_is_half_hessian():
    case = make_case(0.667)
    S = _solver(case, COLLAGEN12)
    rng = np.random.default_rng(1)
    r = np.linalg.norm(case["coords"], axis=1)
    U = rng.normal(size=case["coords"].shape) * (5e-3 * r)[:, None]
    S._prepare()
    S.mesh.displacements[:] = U
    S._update_glo_f_and_K()
    f0, K = S.mesh.forces.copy(), S.K_glo.copy()
    free = np.where(case["movable"])[0]
    for node in free[[3, 40, 200]]:
        eps = 1e-6 * r[node]
        for comp in range(3):
            es, fs = [], []
            for sgn in (1, -1):
                S.mesh.displacements[:] = U
                S.mesh.displacements[node, comp] += sgn * eps
                S._update_glo_f_and_K()
                es.append(S.mesh.strain_energy)
                fs.append(S.mesh.forces.copy())
            dE = (es[0] - es[1]) / (2 * eps)
            assert abs(dE + f0[node, comp]) < 1e-5 * np.abs(f0).max()                    # f = -dE/dU  (A7)
            dF = ((fs[0] - fs[1]) / (2 * eps)).ravel()
            col = -2.0 * K[:, 3 * node + comp].toarray().ravel()                          # K = 1/2 Hessian (A8)
            rows = np.repeat(case["movable"], 3)
            assert np.abs(dF[rows] - col[rows]).max() < 1e-5 * np.abs(col).max()


def test_c_oracle_equals_numpy_oracle():
    case = make_case(0.667)
    S = _solver(case, BATCH_A)
    rng = np.random.default_rng(2)
    U = rng.normal(size=case["coords"].shape) * 2e-7
    U[~case["movable"]] = 0
    S._prepare()
    S.mesh.displacements[:] = U
    e, f, K = S.element_quantities()
    E2, f2, K2 = c_oracle.element_eval(case["coords"], case["tets"], U, S.s, BATCH_A)
    assert np.abs(K - K2).max() < 1e-12 * np.abs(K).max()
    assert np.abs(f - f2).max() < 1e-11 * np.abs(f).max()
    assert np.abs(e - E2).max() < 1e-11 * np.abs(e).max()
    S._update_glo_f_and_K()
    e_tot, forces, Kc = c_oracle.assemble(case["coords"], case["tets"], case["movable"], U, S.s, BATCH_A)
    assert abs(Kc - S.K_glo).max() < 1e-12 * abs(S.K_glo).max()
    assert np.abs(forces - S.mesh.forces).max() < 1e-11 * np.abs(forces).max()
    assert abs(e_tot - S.mesh.strain_energy) < 1e-12 * e_tot


def test_energy_counts_only_tets_with_a_free_node():
    """E_glo leaves out tetrahedra whose four nodes all carry a displacement condition (saenopy `_update_energy`
    [RECALL]).  Invisible under the reference's own boundary conditions (outer shell held at zero displacement,
    simulation.py:127-129), so it is exercised with a non-zero prescribed displacement; Python and C versions agree."""
    case = make_case(0.667, 100.0)
    coords, tets = case["coords"], case["tets"]
    # hold the outer HALF of the domain at a stretched state: many all-fixed tets with non-zero strain
    r = np.linalg.norm(coords, axis=1)
    fixed = r > 0.3 * r.max()
    disp = np.full(coords.shape, np.nan)
    disp[fixed] = 0.01 * coords[fixed]
    forces = np.zeros(coords.shape)
    forces[fixed] = np.nan
    S = Solver()
    S.set_material_model(SemiAffineFiberMaterial(*COLLAGEN12))
    S.set_nodes(coords); S.set_tetrahedra(tets)
    S.set_boundary_condition(disp, forces)
    S._prepare(); S._update_glo_f_and_K()
    e_free = S.mesh.strain_energy
    all_fixed = ~np.any(S.mesh.movable[tets], axis=1)
    assert all_fixed.sum() > 10 and S.mesh.cell_energy[all_fixed].sum() > 0
    assert abs(e_free - S.mesh.cell_energy[~all_fixed].sum()) < 1e-12 * e_free
    S.energy_rule = "all"
    S._update_glo_f_and_K()
    assert abs(S.mesh.strain_energy - S.mesh.cell_energy.sum()) < 1e-12 * e_free and S.mesh.strain_energy > e_free
    e_c, _, _ = c_oracle.assemble(coords, tets, S.mesh.movable, S.mesh.displacements, S.s, COLLAGEN12)
    assert abs(e_c - e_free) < 1e-12 * e_free


def test_relaxation_semantics_and_c_port():
    """Driver semantics (A12-A14): first relrec row logs sum f^2, stop rule fires after i > 6 on the
    last five energies, U moves by step * CG solution; the C port follows the same trajectory."""
    case = make_case(0.667, 100.0)
    S = _solver(case, COLLAGEN12)
    rr = S.solve_boundarycondition(step_size=0.0033, max_iterations=600, rel_conv_crit=0.01)
    assert rr[0, 0] == 0 and len(rr) - 1 == len(S.cg_iterations) and 8 <= len(rr) - 1 < 600
    last5 = rr[-5:, 1]
    assert np.std(last5) / np.sqrt(5) / np.mean(last5) < 0.01
    prev5 = rr[-6:-1, 1]
    assert np.std(prev5) / np.sqrt(5) / np.mean(prev5) >= 0.01
    assert np.all(S.mesh.displacements[~case["movable"]] == 0)
    C = c_oracle.relax(case["coords"], case["tets"], case["movable"], case["f_target"], np.zeros_like(case["coords"]),
                       S.s, COLLAGEN12, 0.0033, 600, 0.01, nthreads=1)
    assert C["n_iter"] == len(rr) - 1
    # two CPU implementations of the same algorithm: CG stops flip by +-1 iteration under rounding,
    # which moves U by ~1e-5 (DESIGN.md "trajectory sensitivity") -- the noise floor of the 1e-4 budget
    assert np.linalg.norm(C["U"] - S.mesh.displacements) / np.linalg.norm(C["U"]) < 1e-4
    assert np.allclose(C["relrec"][:, 1], rr[:, 1], rtol=1e-5)


def test_cg_stop_iteration_is_a_property_of_the_rounding():
    """The same CG on the same matrix stops at a different iteration when only the evaluation of its sums changes
    (plain double as the oracle / fused multiply-add + pairwise sums as a GPU / long double): the iteration count of a
    loosely converged CG is not an invariant two implementations can be held to, the residual criterion is
    (tools/cg_precision_study.py has the full-size numbers quoted in DESIGN.md)."""
    case = make_case(0.3, 10000.0)
    beams = __import__("oracle.saeno_oracle", fromlist=["build_beams"]).build_beams(300)
    r = c_oracle.relax(case["coords"], case["tets"], case["movable"], case["f_target"], np.zeros_like(case["coords"]), beams,
                       COLLAGEN12, 0.0033, 12, 0.01, nthreads=2)
    v = c_oracle.cg_variants(case["coords"], case["tets"], case["movable"], case["f_target"], r["U"], beams, COLLAGEN12)
    counts = [v["double_plain"], v["double_fma_pairwise"], v["long_double"]]
    assert all(300 < c < 3 * len(case["coords"]) for c in counts)
    assert max(counts) - min(counts) > 5              # ~540 / 623 / 487 on this mesh: +-15 % from rounding alone
    assert v["long_double"] <= v["double_plain"]      # the most exact arithmetic converges first


def test_cg_matches_scipy_on_spd_system():
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    rng = np.random.default_rng(0)
    A = sp.random(300, 300, 0.05, random_state=1)
    A = (A @ A.T + 10 * sp.eye(300)).tocsr()
    b = rng.normal(size=300)
    x, it = cg(A, b, maxiter=900, tol=1e-24, return_iterations=True)
    assert np.allclose(x, spl.spsolve(A.tocsc(), b), rtol=1e-8)
    assert np.all(cg(A, np.zeros(300)) == 0)
    x2, it2 = cg(A, b, maxiter=900, tol=1e-5, return_iterations=True)
    r = b - A @ x2
    assert r @ r < 1e-5 * (b @ b) and it2 < it


def test_converged_linear_solution_is_lame():
    """Linear material relaxed with step 1/2 (K is half the Hessian, so step 1/2 is a full Newton
    step) and hard CG convergence reaches the equilibrium; it must be the Lame thick-shell
    solution with E = K_0/6, nu = 1/4 (reference simulation.py:707) up to mesh error."""
    case = make_case(0.2, 0.1)   # 10,261 nodes; the 1,411-node mesh is 13 % too stiff (linear tets)
    r = c_oracle.relax(case["coords"], case["tets"], case["movable"], case["f_target"], np.zeros_like(case["coords"]),
                       build_beams(300), LINEAR, 0.5, 3, 0.0, nthreads=0)
    # the driver uses saenopy's loose CG tolerance; three Newton steps compound it to < 1e-7
    d = np.linalg.norm(case["coords"], axis=1)
    u = np.linalg.norm(r["U"], axis=1)
    sel = (d > 2 * case["r_in"]) & (d < 20 * case["r_in"])
    ratio = u[sel] / lame_shell_displacement(d[sel], 0.1, LINEAR[0], case["r_in"], case["r_out"])
    # measured 0.949: linear tets + a 331-vertex faceted cavity are ~5 % too stiff; what matters is
    # that the deficit is the same at every radius, i.e. the field has the Lame shape B(1/r^2 - r/R^3)
    assert abs(np.median(ratio) - 1.0) < 0.07
    near = ratio[d[sel] < 5 * case["r_in"]]
    far = ratio[d[sel] > 10 * case["r_in"]]
    assert abs(np.median(near) - np.median(far)) < 0.01
    # direction: f_target = +r_hat on the cavity wall means the cavity contracts
    inner = case["bc"]["mask_inner"]
    assert np.all(np.einsum("ij,ij->i", r["U"][inner], case["coords"][inner]) < 0)


# ---- anchors from the reference's shipped tables --------------------------------------------------
def _golden(name):
    path = os.path.join(GOLDEN, name)
    if not os.path.exists(path):
        pytest.skip("%s not generated yet" % name)
    return np.load(path)


def test_reference_table_fixture_shape():
    g = _golden("reference_tables.npz")
    assert np.allclose(g["pressure"], np.logspace(-1, 4, 150), rtol=1e-14)
    x = np.logspace(0, np.log10(50), 101)
    assert np.allclose(g["x"], 10 ** (0.5 * (np.log10(x[1:]) + np.log10(x[:-1]))), rtol=1e-14)  # simulation.py:309-310
    assert g["batchA_y"].shape == (12, 100) and not np.isnan(g["batchA_y"]).any()
    assert np.all(np.diff(g["batchA_y"], axis=0) > 0)      # monotone in pressure at every bin


def _lame_fraction_of_table(y_row, x, pressure, k0):
    sel = (x > 2) & (x < 20)
    exact = lame_shell_displacement(x[sel] * 100e-6, pressure, k0, 100e-6, 10000e-6) / 100e-6
    return float(np.median(y_row[sel] / exact))


def test_shipped_tables_show_early_stopped_iterates():
    """What 'match saenopy' means: the shipped low-pressure rows sit at 0.45x / 0.93x
    of the exact solution, because the reference stops on the five-energy rule long before
    equilibrium.  These two numbers are the anchors the oracle is compared with below."""
    g = _golden("reference_tables.npz")
    assert abs(_lame_fraction_of_table(g["batchA_y"][0], g["x"], 0.1, 1449.0) - 0.4526) < 2e-3
    assert abs(_lame_fraction_of_table(g["linear_y"][0], g["linear_x"], 0.1, 2364.0) - 0.9328) < 2e-3


def test_oracle_reproduces_batchA_early_stop_fraction():
    g = _golden("oracle_batchA_0p1Pa.npz")
    ref = _golden("reference_tables.npz")
    target = _lame_fraction_of_table(ref["batchA_y"][0], ref["x"], 0.1, 1449.0)
    # different (synthetic) mesh => few-% agreement is all these anchors can give
    assert abs(float(g["lame_fraction"]) - target) < 0.06 * target
    # and it is nowhere near the exact-Newton interpretation (K = Hessian would stop at 0.30)
    assert float(g["lame_fraction"]) > 0.38
    sel = (ref["x"] > 1.5) & (ref["x"] < 30) & ~np.isnan(g["curve"])
    ratio = g["curve"][sel] / ref["batchA_y"][0][sel]
    assert abs(np.median(ratio) - 1) < 0.08


def test_oracle_reproduces_linear_table_fraction():
    g = _golden("oracle_linear2364_0p1Pa_step066.npz")
    ref = _golden("reference_tables.npz")
    target = _lame_fraction_of_table(ref["linear_y"][0], ref["linear_x"], 0.1, 2364.0)
    assert abs(float(g["lame_fraction"]) - target) < 0.06 * target


def test_oracle_follows_the_shipped_batchA_table_over_four_decades_of_pressure():
    """The strongest pin available: the reference's own saenopy output (Collagen12_BatchA.npy, recipe
    docs/data/README.md:20-42) against the oracle on the synthetic 10,261-node mesh, at table rows 30..120
    (1 Pa .. 1 kPa: buckling, linear and stiffening regimes; displacements spanning three decades).  Different
    meshes => the agreement is a few per cent, not 1e-3; it corroborates the material law, the half-Hessian
    relaxation and the stop rule at once."""
    g = _golden("oracle_batchA_rows.npz")
    ref = _golden("reference_tables.npz")
    ref_rows = list(ref["rows"])
    x = ref["x"]
    for i, row in enumerate(g["rows"]):
        y_ref = ref["batchA_y"][ref_rows.index(row)]
        y = g["curves"][i]
        ok = ~np.isnan(y) & (x > 1.2) & (x < 40)
        ratio = y[ok] / y_ref[ok]
        if row <= 120:
            assert abs(np.median(ratio) - 1) < 0.06, (row, np.median(ratio))
            assert np.all(np.abs(ratio - 1) < 0.15), (row, ratio.min(), ratio.max())
        else:
            # 10 kPa: a cold start stopped by the five-energy rule after 36 Newton steps is 4-22 % short of the shipped row,
            # increasingly so in the far field -- consistent with the shipped table having been built with a working
            # warm start (get_initial=True with an older saenopy that wrote solver.npz)
            assert 0.75 < np.median(ratio) < 1.0
    assert np.all(np.diff(g["curves"][:, ~np.isnan(g["curves"][0])], axis=0) > 0)
