"""Size-independent properties at BASELINE's larger sizes, where the CPU oracle would take tens of
minutes: Newton's third law (internal forces sum to zero), force = -dE/dU and K = -1/2 df/dU by
directional finite differences, symmetry of K, the streaming CG kernel against its own residual, and
scale invariance of the whole relaxation for the linear material."""
import numpy as np
import pytest

from conftest import COLLAGEN12, LINEAR

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big_case():
    from jointforces_b200.mesh import generate_spherical_inclusion
    from jointforces_b200.simulation import spherical_boundary_conditions
    coords, tets = generate_spherical_inclusion(length_factor=0.03)      # BASELINE configs[3]: 78,057 nodes / 460,172 tets
    assert len(coords) == 78057 and len(tets) == 460172
    bc = spherical_boundary_conditions(coords, 100e-6, 10000e-6, 10000.0)
    movable = np.any(np.isnan(bc["displacement"]), axis=1)
    return dict(coords=coords, tets=tets.astype(np.int32), movable=movable, ft=np.nan_to_num(bc["forces"]))


def _field(case, scale, seed):
    """A smooth field (uniform strain `scale` from a random affine map) plus 0.1 % node-level noise: a rough
    field would put the Delaunay slivers (quality ~1e-3) at strains where exp((lam-ls)/ds) overflows."""
    rng = np.random.default_rng(seed)
    x = case["coords"]
    r = np.linalg.norm(x, axis=1)
    U = scale * (x @ rng.normal(size=(3, 3)).T) + rng.normal(size=x.shape) * (1e-3 * scale * r)[:, None]
    U[~case["movable"]] = 0
    return U


def test_fullsize_force_balance_and_derivatives(big_case, beams):
    from jointforces_b200 import _lib
    c = big_case
    U = _field(c, 1e-2, 0)            # strains of about a percent: all three regimes of collagen12 are populated
    d = _field(c, 1.0, 1)
    d /= np.abs(d).max()
    eps = 1e-6 * np.linalg.norm(c["coords"], axis=1).min()
    with _lib.Context(c["coords"], c["tets"], c["movable"], beams=beams) as ctx:
        assert ctx.info()["cg_resident"] == 0
        ctx.set_material(*COLLAGEN12)
        ctx.set_targets(c["ft"])
        ctx.set_displacements(U); ctx.evaluate()
        f0 = ctx.forces(); e0 = ctx.energy()[0]
        Kd = ctx.apply_stiffness(d)
        KTd_dot = None
        # Newton's third law: every element's nodal forces cancel, so the global sum vanishes
        assert np.abs(f0.sum(0)).max() < 1e-9 * np.abs(f0).sum()
        ctx.set_displacements(U + eps * d); ctx.evaluate()
        fp = ctx.forces(); ep = ctx.energy()[0]
        ctx.set_displacements(U - eps * d); ctx.evaluate()
        fm = ctx.forces(); em = ctx.energy()[0]
        # f = -dE/dU along d
        assert abs((ep - em) / (2 * eps) + np.sum(f0 * d)) < 1e-5 * abs(np.sum(f0 * d))
        # K = -1/2 df/dU (the half Hessian saenopy solves with), on the free rows
        free = c["movable"]
        lhs = (fp - fm)[free] / (2 * eps)
        assert np.linalg.norm(lhs + 2 * Kd[free]) < 1e-5 * np.linalg.norm(2 * Kd[free])
        # symmetry: d2 . K d == d . K d2
        ctx.set_displacements(U); ctx.evaluate()
        d2 = _field(c, 1.0, 2)
        a = np.sum(d2 * ctx.apply_stiffness(d)); b = np.sum(d * ctx.apply_stiffness(d2))
        assert abs(a - b) < 1e-10 * max(abs(a), abs(b))
        # streaming CG: the returned x must satisfy |b - K x|^2 < tol |b|^2 with b = f - f_target
        it = ctx.cg(1e-5)
        x = ctx.cg_solution()
        bvec = (ctx.forces() - c["ft"]) * free[:, None]
        res = bvec - ctx.apply_stiffness(x)
        assert 0 < it[0] < 3 * len(free)
        assert np.sum(res ** 2) < 1.05e-5 * np.sum(bvec ** 2)
        assert e0 > 0


def test_fullsize_relaxation_scales_with_load_for_linear_material(big_case, beams):
    """materials.linear at tiny loads is a linear problem and every stopping test of the algorithm is scale
    invariant, so 2.5x the load must give 2.5x the displacements after the same number of Newton and CG
    iterations -- through element kernel, assembly, streaming CG and the driver at full size."""
    from jointforces_b200 import _lib
    c = big_case
    with _lib.Context(c["coords"], c["tets"], c["movable"], n_sys=2, beams=beams) as ctx:
        ctx.set_material(*LINEAR)
        ctx.set_targets(c["ft"] * 1e-4, 0)
        ctx.set_targets(c["ft"] * 2.5e-4, 1)
        ctx.set_displacements(np.zeros_like(c["coords"]), 0)
        ctx.set_displacements(np.zeros_like(c["coords"]), 1)
        relrec, cgit, n = ctx.relax(0.066, 12, 0.01)
        U0, U1 = ctx.displacements(0), ctx.displacements(1)
    assert n[0] == n[1] == 12
    assert np.abs(cgit[0].astype(int) - cgit[1]).max() <= 0.03 * cgit[0].max()    # CG stop is rounding-sensitive
    assert np.linalg.norm(U1 - 2.5 * U0) < 1e-3 * np.linalg.norm(U1)
    assert np.allclose(relrec[1][:, 1], 6.25 * relrec[0][:, 1], rtol=1e-3)


# ---- committed oracle fixtures at the BASELINE sizes (tests/golden/make_oracle_fullsize.py) -----------------------
import os
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# name -> (kernel the default path must take: (cg_resident, cg_fused), allowed relative difference of the CG iteration total).
# The CUDA kernels need 0.6-3 % FEWER iterations than the oracle's C port, the classic two-reduction kernels exactly like
# the single-reduction ones (tools/cg_iter_bias.py): CG's convergence is delayed by rounding, and the kernels round
# less (fused multiply-adds, tree sums, exact all-reduce).  The oracle's own CG on its own matrix shows the same shift when
# only its sums are evaluated that way, and more in long double (tools/cg_precision_study.py: 1005 / 984 / 946 iterations
# at 10 kPa, 825 / 807 / 804 at 100 Pa); the displacements agree to 1e-5 regardless.  Hence an asymmetric window.
FULL_CASES = {
    "cfg2_full": ((2, 1), (-0.04, 0.01)),   # register-resident single-reduction kernel (the bench headline)
    "cfg3_rowA": ((2, 1), (-0.04, 0.01)),
    "cfg3_rowB": ((2, 1), (-0.05, 0.01)),
    "cfg4_full": ((0, 2), (-0.05, 0.01)),   # streaming single-reduction kernel
    "cfg5_loose": ((0, 0), (-0.05, 0.01)),  # plain HBM-bound streaming kernel (roofline_large of bench.py)
    "cfg5_tight": ((0, 0), (-0.05, 0.01)),
}


def _full_case(g):
    from jointforces_b200.mesh import generate_spherical_inclusion
    from jointforces_b200.simulation import spherical_boundary_conditions
    coords, tets = generate_spherical_inclusion(length_factor=float(g["length_factor"]))
    assert len(coords) == int(g["n_nodes"]) and len(tets) == int(g["n_tets"])
    bc = spherical_boundary_conditions(coords, 100e-6, 10000e-6, float(g["pressure"]))
    movable = np.any(np.isnan(bc["displacement"]), axis=1)
    return coords, tets.astype(np.int32), movable, np.nan_to_num(bc["forces"])


@pytest.mark.parametrize("name", list(FULL_CASES))
def test_fullsize_relaxation_matches_oracle_fixture(name, beams, capsys):
    """The CUDA path against the committed CPU-oracle solution on the mesh the benchmark numbers are quoted on:
    Newton iterations equal, displacements within 1e-4 relative L2 (on the stored node subset), lookup-table curve
    within 1e-3 (BASELINE tolerances); the signed mean CG-iteration difference per Newton step is reported."""
    from jointforces_b200 import _lib
    from jointforces_b200.simulation import radial_median_curve, lookup_bins
    path = os.path.join(GOLDEN, "oracle_full_%s.npz" % name)
    if not os.path.exists(path):
        pytest.skip("fixture missing")
    g = np.load(path)
    kernel, cg_tol_total = FULL_CASES[name]
    coords, tets, movable, ft = _full_case(g)
    with _lib.Context(coords, tets, movable, beams=beams) as ctx:
        info = ctx.info()
        ctx.set_material(*g["material"])
        ctx.set_targets(ft)
        ctx.set_displacements(np.zeros_like(coords))
        relrec, cgit, n = ctx.relax(float(g["step"]), int(g["max_iter"]), 0.01, cg_tol=float(g["cg_tol"]))
        U = ctx.displacements()
    fused = info["cg_fused"] if float(g["cg_tol"]) == 0.0 else 0   # tight tolerances take the classic two-reduction kernels
    assert (info["cg_resident"], info["cg_fused"]) == kernel, info
    assert int(n[0]) == len(g["relrec"]) - 1
    stride = int(g["stride"])
    err = np.linalg.norm(U[::stride] - g["U_sub"]) / np.linalg.norm(g["U_sub"])
    d = cgit[0].astype(np.int64) - g["cg_iters"]
    with capsys.disabled():
        print("\n[%s] %d nodes, kernel (resident %d, fused %d), %d Newton steps, CG iterations %d vs oracle %d: signed mean "
              "difference %+.2f per step (max |d| %d), |U-U_oracle|/|U_oracle| = %.2e"
              % (name, len(coords), info["cg_resident"], fused, n[0], cgit[0].sum(), g["cg_iters"].sum(), d.mean(),
                 np.abs(d).max(), err))
    assert err < 1e-4, err
    assert abs(np.linalg.norm(U) / float(g["U_norm"]) - 1) < 1e-4
    assert np.allclose(relrec[0][:, 1], g["relrec"][:, 1], rtol=1e-4)
    assert cg_tol_total[0] <= d.sum() / g["cg_iters"].sum() <= cg_tol_total[1]
    x, _ = lookup_bins()
    curve = radial_median_curve(coords, U, 100e-6, x)
    ok = ~np.isnan(g["curve"])
    assert np.array_equal(np.isnan(curve), ~ok)
    assert np.max(np.abs(curve[ok] / g["curve"][ok] - 1)) < 1e-3


def test_two_load_cases_side_by_side_match_their_oracle_fixtures(beams, capsys):
    """The sweeps' kernel on the configs[2] mesh: a context of TWO systems runs both CG solves in one launch, two CTAs
    per SM, each system its own reduction group (jf_cg_pair_kernel, info cg_fused == 4).  Rows A and B of the
    lookup-table sweep (different pressures: 95 and 53 Newton steps, different CG iteration counts in every step)
    relaxed together must each match their own oracle fixture exactly as the single solves do."""
    from jointforces_b200 import _lib
    from jointforces_b200.simulation import radial_median_curve, lookup_bins
    paths = [os.path.join(GOLDEN, "oracle_full_cfg3_row%s.npz" % r) for r in "AB"]
    if not all(os.path.exists(p) for p in paths):
        pytest.skip("fixture missing")
    gs = [np.load(p) for p in paths]
    cases = [_full_case(g) for g in gs]
    coords, tets, movable, _ = cases[0]
    with _lib.Context(coords, tets, movable, n_sys=2, beams=beams) as ctx:
        info = ctx.info()
        assert info["cg_fused"] == 4, info
        ctx.set_material(*gs[0]["material"])
        for s, c in enumerate(cases):
            ctx.set_targets(c[3], s)
            ctx.set_displacements(np.zeros_like(coords), s)
        relrec, cgit, n = ctx.relax(float(gs[0]["step"]), int(gs[0]["max_iter"]), 0.01)
        U = [ctx.displacements(s) for s in range(2)]
        st = ctx.stats()
    x, _ = lookup_bins()
    for s, g in enumerate(gs):
        assert int(n[s]) == len(g["relrec"]) - 1
        stride = int(g["stride"])
        err = np.linalg.norm(U[s][::stride] - g["U_sub"]) / np.linalg.norm(g["U_sub"])
        d = cgit[s][:int(n[s])].astype(np.int64) - g["cg_iters"]
        with capsys.disabled():
            print("\n[pair, system %d] %d Newton steps, CG iterations %d vs oracle %d: signed mean difference %+.2f per step, "
                  "|U-U_oracle|/|U_oracle| = %.2e" % (s, n[s], cgit[s].sum(), g["cg_iters"].sum(), d.mean(), err))
        assert err < 1e-4, err
        assert np.allclose(relrec[s][:int(n[s]) + 1, 1], g["relrec"][:, 1], rtol=1e-4)
        assert -0.05 <= d.sum() / g["cg_iters"].sum() <= 0.01
        curve = radial_median_curve(coords, U[s], 100e-6, x)
        ok = ~np.isnan(g["curve"])
        assert np.max(np.abs(curve[ok] / g["curve"][ok] - 1)) < 1e-3
    with capsys.disabled():
        print("[pair] device time %.1f ms for both (CG %.1f ms, %d iterations in total)" % (st["ms_relax"], st["ms_cg"], st["cg_iterations"]))


def test_pair_kernel_is_bitwise_reproducible_and_independent_of_its_neighbour(beams):
    """Integer-accumulated sums and per-system reduction groups: a system's trajectory does not depend on what the other
    system of the launch is doing, nor on the run."""
    from jointforces_b200 import _lib
    g = np.load(os.path.join(GOLDEN, "oracle_full_cfg3_rowA.npz"))
    coords, tets, movable, ft = _full_case(g)
    out = []
    with _lib.Context(coords, tets, movable, n_sys=2, beams=beams) as ctx:
        assert ctx.info()["cg_fused"] == 4
        ctx.set_material(*g["material"])
        for other in (1.0, 37.0, 37.0):
            ctx.set_targets(ft, 0)
            ctx.set_targets(ft * other, 1)
            for s in range(2):
                ctx.set_displacements(np.zeros_like(coords), s)
            relrec, cgit, n = ctx.relax(float(g["step"]), 12, 0.01)
            out.append((ctx.displacements(0), cgit[0].copy(), ctx.displacements(1)))
    assert np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][0], out[1][0])   # neighbour changed: same bits
    assert np.array_equal(out[1][0], out[2][0]) and np.array_equal(out[1][2], out[2][2])   # same run twice: same bits
    assert np.array_equal(out[0][0], out[0][2])                                            # same load case in both slots: same bits
