"""GPU parity: the CUDA path through the C ABI against the CPU oracle on the same inputs."""
import numpy as np
import pytest

from conftest import COLLAGEN12, LINEAR, FIBRIN, BATCH_A, make_case

pytestmark = pytest.mark.gpu


def _random_U(case, scale, seed=0):
    rng = np.random.default_rng(seed)
    r = np.linalg.norm(case["coords"], axis=1)
    U = rng.normal(size=case["coords"].shape) * (scale * r)[:, None]
    U[~case["movable"]] = 0.0
    return U


@pytest.mark.parametrize("mat", [COLLAGEN12, LINEAR, FIBRIN, BATCH_A])
@pytest.mark.parametrize("scale", [1e-7, 3e-3, 5e-2])
def test_element_quantities_match_oracle(mat, scale, beams):
    from jointforces_b200 import _lib
    from oracle import c_oracle
    case = make_case(0.667)
    U = _random_U(case, scale)
    E0, f0, K0 = c_oracle.element_eval(case["coords"], case["tets"], U, beams, mat)
    with _lib.Context(case["coords"], case["tets"], case["movable"], beams=beams) as ctx:
        ctx.set_material(*mat)
        ctx.set_displacements(U)
        ctx.evaluate()
        E, f, K = ctx.element_quantities()
    # per-tet comparison.  K and E: relative to the tet's largest entry.  f carries the rounding of
    # lambda = |F s| - 1 (absolute ~1e-16 in BOTH implementations), i.e. an absolute noise floor
    # of ~1e-16 * |K_t| * h_t on top of the relative error.
    def rel(a, b):
        a = a.reshape(len(a), -1); b = b.reshape(len(b), -1)
        return np.max(np.abs(a - b).max(1) / np.maximum(np.abs(b).max(1), 1e-300))
    assert rel(K, K0) < 1e-11
    B = case["coords"][case["tets"][:, 1:]] - case["coords"][case["tets"][:, :1]]
    h = np.abs(np.linalg.det(B)) ** (1.0 / 3.0)
    tol = 1e-10 * np.abs(f0).reshape(len(f0), -1).max(1) + 1e-13 * np.abs(K0).reshape(len(K0), -1).max(1) * h
    assert np.all(np.abs(f - f0).reshape(len(f0), -1).max(1) <= tol)
    # E ~ lambda^2: the same 1e-16 absolute noise in lambda is 2e-16/strain relative in E
    assert np.abs(E - E0).max() <= (1e-10 + 1e-15 / scale) * np.abs(E0).max() + 1e-300


@pytest.mark.parametrize("mat", [COLLAGEN12, LINEAR])
def test_assembly_matches_oracle(mat, beams):
    from jointforces_b200 import _lib
    from oracle import c_oracle
    case = make_case(0.667)
    U = _random_U(case, 1e-2, seed=3)
    e0, f0, K0 = c_oracle.assemble(case["coords"], case["tets"], case["movable"], U, beams, mat)
    with _lib.Context(case["coords"], case["tets"], case["movable"], beams=beams) as ctx:
        ctx.set_material(*mat)
        ctx.set_targets(case["f_target"])
        ctx.set_displacements(U)
        ctx.evaluate()
        e, r2, f2 = ctx.energy()
        f = ctx.forces()
        K = ctx.stiffness_csr()
    assert abs(e - e0) < 1e-11 * abs(e0)
    assert np.abs(f - f0).max() < 1e-10 * np.abs(f0).max()
    # the library drops the columns of fixed nodes (they multiply zeros in CG); compare free x free
    free = np.repeat(case["movable"], 3)
    D = (K - K0)[free][:, free]
    assert abs(D).max() < 1e-11 * abs(K0).max()
    assert K[~free].nnz == 0
    m = case["movable"]
    assert abs(f2 - np.sum(f0[m] ** 2)) < 1e-10 * f2
    assert abs(r2 - np.sum((f0[m] - case["f_target"][m]) ** 2)) < 1e-10 * r2


def test_cg_matches_oracle(beams):
    from jointforces_b200 import _lib
    from oracle import c_oracle
    from oracle.saeno_oracle import cg
    case = make_case(0.667)
    U = _random_U(case, 1e-3, seed=5)
    _, f0, K0 = c_oracle.assemble(case["coords"], case["tets"], case["movable"], U, beams, COLLAGEN12)
    b = f0 - case["f_target"]
    b[~case["movable"]] = 0
    x0, it0 = cg(K0, b.ravel(), maxiter=3 * len(b), tol=1e-5, return_iterations=True)
    x1, it1 = cg(K0, b.ravel(), maxiter=3 * len(b), tol=1e-20, return_iterations=True)
    with _lib.Context(case["coords"], case["tets"], case["movable"], beams=beams) as ctx:
        ctx.set_material(*COLLAGEN12)
        ctx.set_targets(case["f_target"])
        ctx.set_displacements(U)
        ctx.evaluate()
        it = ctx.cg(1e-5)
        x = ctx.cg_solution()
        it_t = ctx.cg(1e-20)
        x_t = ctx.cg_solution()
    assert abs(int(it[0]) - it0) <= 1
    assert np.linalg.norm(x.ravel() - x0) < 5e-3 * np.linalg.norm(x0)
    assert np.linalg.norm(x_t.ravel() - x1) < 1e-7 * np.linalg.norm(x1)
    assert np.all(x[~case["movable"]] == 0)


@pytest.mark.parametrize("mat,pressure,step", [(COLLAGEN12, 100.0, 0.0033), (LINEAR, 1.0, 0.066), (BATCH_A, 1000.0, 0.0033)])
def test_relaxation_matches_oracle(mat, pressure, step, beams):
    """The BASELINE tolerance: relaxed displacements within 1e-4 relative L2 of the oracle."""
    from jointforces_b200 import _lib
    from oracle import c_oracle
    case = make_case(0.667, pressure)
    ref = c_oracle.relax(case["coords"], case["tets"], case["movable"], case["f_target"], np.zeros_like(case["coords"]),
                         beams, mat, step, 600, 0.01, nthreads=1)
    out = _lib.solve_boundarycondition(case["coords"], case["tets"], case["movable"], case["f_target"],
                                       np.zeros_like(case["coords"]), mat, step, 600, 0.01, beams=beams)
    assert out["n_iter"] == ref["n_iter"]
    err = np.linalg.norm(out["U"] - ref["U"]) / np.linalg.norm(ref["U"])
    assert err < 1e-4, err
    assert np.allclose(out["relrec"][:, 1], ref["relrec"][:, 1], rtol=1e-4)
    # CG stops flip by a few iterations under rounding (two CPU builds of the oracle differ likewise)
    assert np.abs(out["cg_iters"].astype(int) - ref["cg_iters"]).max() <= 12
    assert abs(int(out["cg_iters"].sum()) - int(ref["cg_iters"].sum())) <= 0.02 * ref["cg_iters"].sum()
    assert np.abs(out["forces"] - ref["forces"]).max() < 1e-3 * np.abs(ref["forces"]).max()


def test_relaxation_tight_cg_matches_oracle_closely(beams):
    """With CG converged hard the iteration-count sensitivity disappears: agreement ~1e-9."""
    from jointforces_b200 import _lib
    from oracle.saeno_oracle import Solver, SemiAffineFiberMaterial, cg
    import oracle.saeno_oracle as so
    case = make_case(0.667, 100.0)
    S = Solver()
    S.set_material_model(SemiAffineFiberMaterial(*COLLAGEN12))
    S.set_nodes(case["coords"]); S.set_tetrahedra(case["tets"])
    S.set_boundary_condition(case["bc"]["displacement"], case["bc"]["forces"])
    S._prepare(); S._update_glo_f_and_K()
    m = S.mesh
    for _ in range(5):
        ff = m.forces - m.forces_target
        ff[~m.movable] = 0
        uu = cg(S.K_glo, ff.ravel(), maxiter=3 * len(ff), tol=1e-22).reshape(ff.shape)
        m.displacements[m.movable] += 0.05 * uu[m.movable]
        S._update_glo_f_and_K()
    with _lib.Context(case["coords"], case["tets"], case["movable"], beams=beams) as ctx:
        ctx.set_material(*COLLAGEN12)
        ctx.set_targets(case["f_target"])
        ctx.set_displacements(np.zeros_like(case["coords"]))
        relrec, cgit, n = ctx.relax(0.05, 5, 0.0, cg_tol=1e-22)
        U = ctx.displacements()
    assert n[0] == 5
    assert np.linalg.norm(U - m.displacements) / np.linalg.norm(U) < 1e-8


def test_lockstep_batch_equals_single_solves(beams, monkeypatch):
    """Three load cases in one lockstep batch against three single solves.  With the same reduction (slot tables,
    JF_CG_SLOT_SUMS=1) the arithmetic is the same per system: identical bits.  By default a single system reduces
    its CG sums through the fixed-point accumulators (exact sums, one rounding), a batch through the slot tables
    (sequential double sums): agreement to the trajectory noise floor (DESIGN.md "Trajectory sensitivity")."""
    from jointforces_b200 import _lib
    pressures = [1.0, 100.0, 3000.0]
    cases = [make_case(0.667, p) for p in pressures]
    c0 = cases[0]

    def run_singles():
        return [_lib.solve_boundarycondition(c["coords"], c["tets"], c["movable"], c["f_target"], np.zeros_like(c["coords"]),
                                             COLLAGEN12, 0.0033, 600, 0.01, beams=beams) for c in cases]
    singles = run_singles()
    monkeypatch.setenv("JF_CG_SLOT_SUMS", "1")
    singles_slot = run_singles()
    monkeypatch.delenv("JF_CG_SLOT_SUMS")
    with _lib.Context(c0["coords"], c0["tets"], c0["movable"], n_sys=3, beams=beams) as ctx:
        ctx.set_material(*COLLAGEN12)
        for s, c in enumerate(cases):
            ctx.set_targets(c["f_target"], s)
            ctx.set_displacements(np.zeros_like(c["coords"]), s)
        relrec, cgit, n = ctx.relax(0.0033, 600, 0.01)
        for s in range(3):
            assert n[s] == singles_slot[s]["n_iter"]
            assert np.array_equal(ctx.displacements(s), singles_slot[s]["U"])  # same kernels, same order: bitwise
            assert np.array_equal(relrec[s], singles_slot[s]["relrec"])
            assert n[s] == singles[s]["n_iter"]
            # noise floor of this 360-node case: the oracle on 8 threads differs from the oracle on 1 thread by 1.1e-4
            # at 1 Pa (3.5e-5 at 100 Pa, 1e-5 at 3 kPa), the kernel variants from each other and from the oracle by
            # 0.2-1.8e-4 / 0.5-3e-5 / 0.3-2e-5 (tools/trajectory_noise.py); the BASELINE-size meshes sit at 1e-5
            assert np.linalg.norm(ctx.displacements(s) - singles[s]["U"]) / np.linalg.norm(singles[s]["U"]) < 3e-4


def test_bitwise_reproducible(beams):
    from jointforces_b200 import _lib
    c = make_case(0.667, 300.0)
    a = _lib.solve_boundarycondition(c["coords"], c["tets"], c["movable"], c["f_target"], np.zeros_like(c["coords"]),
                                     COLLAGEN12, 0.0033, 40, 0.01, beams=beams)
    b = _lib.solve_boundarycondition(c["coords"], c["tets"], c["movable"], c["f_target"], np.zeros_like(c["coords"]),
                                     COLLAGEN12, 0.0033, 40, 0.01, beams=beams)
    assert np.array_equal(a["U"], b["U"]) and np.array_equal(a["relrec"], b["relrec"])


def test_reordering_and_folding_are_transparent(beams):
    """Internal renumbering / antipodal folding must not change results beyond rounding."""
    from jointforces_b200 import _lib
    c = make_case(0.667, 100.0)
    U = _random_U(c, 1e-2, seed=9)
    outs = []
    for flags in (0, _lib.JF_FLAG_NO_REORDER | _lib.JF_FLAG_NO_FOLD):
        with _lib.Context(c["coords"], c["tets"], c["movable"], beams=beams, flags=flags) as ctx:
            ctx.set_material(*COLLAGEN12)
            ctx.set_targets(c["f_target"])
            ctx.set_displacements(U)
            ctx.evaluate()
            outs.append((ctx.forces(), ctx.stiffness_csr(), ctx.energy()[0], ctx.info()))
    assert outs[0][3]["n_dirs"] == 172 and outs[1][3]["n_dirs"] == 286
    assert np.abs(outs[0][0] - outs[1][0]).max() < 1e-11 * np.abs(outs[1][0]).max()
    assert abs(outs[0][1] - outs[1][1]).max() < 1e-11 * abs(outs[1][1]).max()
    assert abs(outs[0][2] - outs[1][2]) < 1e-12 * abs(outs[1][2])


# ---- committed oracle fixtures at BASELINE configs[0] size (10,261 nodes / 59,520 tets) ----------------
import os
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["cfg1_linear_100Pa", "batchA_0p1Pa", "linear2364_0p1Pa_step066", "collagen12_100Pa"])
def test_golden_oracle_solutions(name, beams):
    """GPU relaxation vs the committed oracle solution (tests/golden/make_oracle_golden.py):
    displacements within 1e-4 relative L2, lookup-table curve within 1e-3 (BASELINE tolerances)."""
    from jointforces_b200 import _lib
    from jointforces_b200.simulation import radial_median_curve, lookup_bins
    path = os.path.join(GOLDEN, "oracle_%s.npz" % name)
    if not os.path.exists(path):
        pytest.skip("fixture missing")
    g = np.load(path)
    case = make_case(float(g["lf_iso"]), float(g["pressure"]))
    assert len(case["coords"]) == int(g["n_nodes"]) and len(case["tets"]) == int(g["n_tets"])
    out = _lib.solve_boundarycondition(case["coords"], case["tets"], case["movable"], case["f_target"],
                                       np.zeros_like(case["coords"]), tuple(g["material"]), float(g["step"]), 600, 0.01,
                                       beams=beams)
    assert out["n_iter"] == len(g["relrec"]) - 1
    err = np.linalg.norm(out["U"] - g["U"]) / np.linalg.norm(g["U"])
    assert err < 1e-4, err
    x, _ = lookup_bins()
    curve = radial_median_curve(case["coords"], out["U"], case["r_in"], x)
    ok = ~np.isnan(g["curve"])
    assert np.array_equal(np.isnan(curve), ~ok)
    assert np.max(np.abs(curve[ok] / g["curve"][ok] - 1)) < 1e-3
    assert np.allclose(out["relrec"][:, 1], g["relrec"][:, 1], rtol=1e-4)


def test_resident_and_streaming_cg_agree(beams, monkeypatch):
    """Same arithmetic, different data placement (matrix in registers / in shared memory / streamed
    from global memory): the three CG kernels must follow the same trajectory."""
    from jointforces_b200 import _lib
    c = make_case(0.3, 100.0)
    res = []
    NR = _lib.JF_FLAG_NO_RESIDENT
    for flags, no_regres, classic in ((0, False, False), (0, False, True), (0, True, True), (NR, False, False), (NR, False, True)):
        for name, on in (("JF_CG_NO_REGRES", no_regres), ("JF_CG_CLASSIC", classic)):
            if on:
                monkeypatch.setenv(name, "1")
            else:
                monkeypatch.delenv(name, raising=False)
        with _lib.Context(c["coords"], c["tets"], c["movable"], beams=beams, flags=flags) as ctx:
            ctx.set_material(*COLLAGEN12)
            ctx.set_targets(c["f_target"])
            ctx.set_displacements(np.zeros_like(c["coords"]))
            relrec, cgit, n = ctx.relax(0.0033, 25, 0.01)
            info = ctx.info()
            res.append((ctx.displacements(), cgit[0], (info["cg_resident"], info["cg_fused"])))
    kinds = [r[2] for r in res]
    # single-reduction register kernel / classic register kernel / classic shared-memory kernel /
    # single-reduction streaming kernel / classic streaming kernel
    assert kinds == [(2, 1), (2, 0), (1, 0), (0, 2), (0, 0)], kinds
    for other in res[:4]:
        # CG's loose stop makes the iteration count rounding-sensitive (DESIGN.md "Trajectory sensitivity")
        assert np.abs(other[1].astype(int) - res[4][1]).max() <= 0.05 * res[4][1].max()
        assert np.linalg.norm(other[0] - res[4][0]) / np.linalg.norm(res[4][0]) < 5e-5


@pytest.mark.parametrize("streamed", [False, True])
def test_fixed_point_reduction_falls_back_when_a_partial_does_not_fit(beams, monkeypatch, streamed):
    """JF_FX_TEST_OVERFLOW=3 makes CTA 0 declare every third fixed-point reduction unrepresentable: all CTAs see the
    mark in the accumulator sector and redo that iteration's sums through the slot tables.  Same trajectory as the
    pure slot-table run up to the noise floor, and no deadlock."""
    from jointforces_b200 import _lib
    c = make_case(0.3, 100.0)
    res = {}
    for name, env in (("fixed", {}), ("mixed", {"JF_FX_TEST_OVERFLOW": "3"}), ("slots", {"JF_CG_SLOT_SUMS": "1"})):
        for k in ("JF_FX_TEST_OVERFLOW", "JF_CG_SLOT_SUMS"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        with _lib.Context(c["coords"], c["tets"], c["movable"], beams=beams, flags=_lib.JF_FLAG_NO_RESIDENT if streamed else 0) as ctx:
            assert ctx.info()["cg_fused"] == (2 if streamed else 1)
            ctx.set_material(*COLLAGEN12)
            ctx.set_targets(c["f_target"])
            ctx.set_displacements(np.zeros_like(c["coords"]))
            relrec, cgit, n = ctx.relax(0.0033, 600, 0.01)
            res[name] = (ctx.displacements(), cgit[0].copy(), int(n[0]))
    assert res["fixed"][2] == res["mixed"][2] == res["slots"][2]
    for other in ("mixed", "slots"):
        assert np.abs(res[other][1].astype(int) - res["fixed"][1]).max() <= 0.10 * res["fixed"][1].max()   # CG stop is rounding-sensitive
        assert np.linalg.norm(res[other][0] - res["fixed"][0]) / np.linalg.norm(res["fixed"][0]) < 5e-5
    assert not np.array_equal(res["mixed"][0], res["fixed"][0])          # the mark really changed the path taken


def test_device_and_host_structure_build_agree(beams, monkeypatch):
    """Adjacency, ordering scores, shape tensors and gather lists are built on the device; JF_HOST_SETUP=1 keeps the
    adjacency and the scores on the host.  Same structure -> same arithmetic -> identical bits."""
    from jointforces_b200 import _lib
    c = make_case(0.3, 100.0)
    res = []
    for host in (False, True):
        if host:
            monkeypatch.setenv("JF_HOST_SETUP", "1")
        else:
            monkeypatch.delenv("JF_HOST_SETUP", raising=False)
        with _lib.Context(c["coords"], c["tets"], c["movable"], beams=beams) as ctx:
            ctx.set_material(*COLLAGEN12)
            ctx.set_targets(c["f_target"])
            ctx.set_displacements(np.zeros_like(c["coords"]))
            relrec, cgit, n = ctx.relax(0.0033, 12, 0.01)
            res.append((ctx.displacements(), cgit[0].copy(), ctx.info()))
    assert res[0][2]["n_slots"] == res[1][2]["n_slots"] and res[0][2]["n_blocks"] == res[1][2]["n_blocks"]
    assert np.array_equal(res[0][1], res[1][1])
    assert np.array_equal(res[0][0], res[1][0])


@pytest.mark.parametrize("streamed", [False, True])
def test_batch_of_two_systems(beams, streamed, monkeypatch):
    """Two load cases in lockstep: small enough to be resident together (one CTA group per system), and through
    the streaming kernels (systems finish their CG solves at different iterations)."""
    from jointforces_b200 import _lib
    cases = [make_case(0.667, p) for p in (10.0, 1000.0)]

    def run_singles():
        return [_lib.solve_boundarycondition(c["coords"], c["tets"], c["movable"], c["f_target"], np.zeros_like(c["coords"]),
                                             COLLAGEN12, 0.0033, 30, 0.01, beams=beams) for c in cases]
    monkeypatch.setenv("JF_CG_SLOT_SUMS", "1")       # the batch reduces through the slot tables: compare like with like
    singles = run_singles()
    monkeypatch.delenv("JF_CG_SLOT_SUMS")
    singles_fx = run_singles()                        # default: fixed-point accumulators (exact sums)
    c0 = cases[0]
    with _lib.Context(c0["coords"], c0["tets"], c0["movable"], n_sys=2, beams=beams,
                      flags=_lib.JF_FLAG_NO_RESIDENT if streamed else 0) as ctx:
        info = ctx.info()
        assert (info["cg_resident"], info["cg_fused"]) == ((0, 2) if streamed else (2, 1)), info
        ctx.set_material(*COLLAGEN12)
        for s, c in enumerate(cases):
            ctx.set_targets(c["f_target"], s)
            ctx.set_displacements(np.zeros_like(c["coords"]), s)
        relrec, cgit, n = ctx.relax(0.0033, 30, 0.01)
        for s in range(2):
            # the grid (hence the summation order of the dot products) differs from the single-system
            # launch, so agreement is to rounding-amplified-by-CG level, not bitwise
            # (the streamed run also differs from the singles in kernel: CG's loose stop makes the iteration count
            # rounding-sensitive, DESIGN.md "Trajectory sensitivity")
            assert np.abs(cgit[s].astype(int) - singles[s]["cg_iters"]).max() <= (0.15 * singles[s]["cg_iters"].max() if streamed else 2)
            assert np.linalg.norm(ctx.displacements(s) - singles[s]["U"]) / np.linalg.norm(singles[s]["U"]) < (5e-4 if streamed else 2e-5)   # streamed: another kernel, 30 unconverged steps (noise floor below)
            # thirty early-stopped, unconverged steps after a differently rounded reduction: the noise floor of such
            # iterates is a few 1e-4 (1.7e-4 measured); it shrinks as the relaxation converges (5e-5 at the stop rule,
            # test_lockstep_batch_equals_single_solves)
            assert np.linalg.norm(ctx.displacements(s) - singles_fx[s]["U"]) / np.linalg.norm(singles_fx[s]["U"]) < 5e-4


# ---- edge cases and error behaviour -----------------------------------------------------------------------
def test_error_paths(beams):
    from jointforces_b200 import _lib
    nodes = np.array([[0., 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 1]])
    good = np.array([[0, 1, 2, 3], [1, 2, 3, 4]], np.int32)
    mov = np.array([1, 1, 1, 1, 0], np.uint8)
    with pytest.raises(_lib.JFError, match="outside"):
        _lib.Context(nodes, np.array([[0, 1, 2, 7]], np.int32), mov)
    flat = nodes.copy(); flat[:, 2] = 0
    with pytest.raises(_lib.JFError, match="degenerate"):
        _lib.Context(flat, good, mov)
    with pytest.raises(_lib.JFError, match="n_sys"):
        _lib.Context(nodes, good, mov, n_sys=1000)
    with _lib.Context(nodes, good, mov, beams=beams) as ctx:
        with pytest.raises(_lib.JFError, match="material"):
            ctx.evaluate()
        ctx.set_material(100.0, 1e30, 1e30, 1e30)
        with pytest.raises(_lib.JFError, match="targets"):
            ctx.relax(0.1, 3, 0.01)
        with pytest.raises(_lib.JFError, match="before"):
            ctx.forces()
        with pytest.raises(_lib.JFError, match="system index"):
            ctx.set_targets(np.zeros((5, 3)), sys=3)
        with pytest.raises(_lib.JFError):
            ctx.set_material(-1.0, 1.0, 1.0, 1.0)


def test_degenerate_problems_do_not_crash(beams):
    """All nodes fixed (no free rows), an orphan free node (empty stiffness row), zero load."""
    from jointforces_b200 import _lib
    nodes = np.array([[0., 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 1], [5, 5, 5]])
    tets = np.array([[0, 1, 2, 3], [1, 2, 3, 4]], np.int32)
    # all fixed
    with _lib.Context(nodes, tets, np.zeros(6, np.uint8), beams=beams) as ctx:
        ctx.set_material(100.0, 1e30, 1e30, 1e30)
        ctx.set_targets(np.zeros((6, 3)))
        U = np.zeros((6, 3)); U[4] = [0, 0, 0.01]
        ctx.set_displacements(U)
        relrec, cgit, n = ctx.relax(0.1, 3, 0.01)
        assert np.array_equal(ctx.displacements(), U) and np.all(cgit[0] == 0)
        # the prescribed stretch strains tet 1, but E_glo counts only tets with a free node (saenopy `_update_energy`,
        # oracle `energy_rule`): none here.  The per-tet energies keep it.
        assert relrec[0][-1, 1] == 0
        assert ctx.element_quantities(want_K=False)[0][1] > 0
    # node 5 is free but belongs to no tet (empty stiffness row); nodes 1-4 fixed so the problem is well posed
    mov = np.array([1, 0, 0, 0, 0, 1], np.uint8)
    with _lib.Context(nodes, tets, mov, beams=beams) as ctx:
        ctx.set_material(100.0, 1e30, 1e30, 1e30)
        ctx.set_targets(np.zeros((6, 3)))
        ctx.set_displacements(np.zeros((6, 3)))
        relrec, cgit, n = ctx.relax(0.1, 2, 0.0)
        # zero load: only the rounding of |F s| - 1 (1e-16, same in the oracle) is left to "relax"
        assert np.abs(ctx.displacements()).max() < 1e-12 and np.all(relrec[0][:, 1] < 1e-25)
        # and with a load on a connected node the orphan stays put while the others move
        ft = np.zeros((6, 3)); ft[0] = [1e-3, 0, 0]
        ctx.set_targets(ft)
        ctx.set_displacements(np.zeros((6, 3)))
        relrec, cgit, n = ctx.relax(0.5, 2, 0.0)
        Uo = ctx.displacements()
        assert np.all(Uo[5] == 0) and np.all(Uo[1:5] == 0) and 1e-8 < np.abs(Uo[0]).max() < 1e-2 and np.isfinite(Uo).all()


def test_matches_oracle_on_structured_mesh_and_other_materials(beams):
    """icoshell mesh (different connectivity pattern), fibrin (buckling only), high pressure (stiffening regime)."""
    from jointforces_b200 import _lib
    from oracle import c_oracle
    for mat, p in ((FIBRIN, 30.0), (COLLAGEN12, 5000.0)):
        case = make_case(0.667, p, method="icoshell")
        ref = c_oracle.relax(case["coords"], case["tets"], case["movable"], case["f_target"], np.zeros_like(case["coords"]),
                             beams, mat, 0.0033, 600, 0.01, nthreads=1)
        out = _lib.solve_boundarycondition(case["coords"], case["tets"], case["movable"], case["f_target"],
                                           np.zeros_like(case["coords"]), mat, 0.0033, 600, 0.01, beams=beams)
        assert out["n_iter"] == ref["n_iter"]
        assert np.linalg.norm(out["U"] - ref["U"]) / np.linalg.norm(ref["U"]) < 1e-4


def test_irregular_random_mesh_matches_oracle(beams):
    """Unstructured input: random points, Delaunay tets of wildly varying quality and node degree (rows wider than
    the register-resident kernel takes), arbitrary fixed/loaded node sets, non-zero prescribed displacements."""
    from scipy.spatial import Delaunay
    from jointforces_b200 import _lib
    from oracle import c_oracle
    rng = np.random.default_rng(11)
    pts = rng.random((500, 3)) * 1e-3
    tets = Delaunay(pts).simplices.astype(np.int32)
    B = pts[tets[:, 1:]] - pts[tets[:, :1]]
    vol = np.abs(np.linalg.det(B)) / 6
    tets = tets[vol > 1e-4 * np.median(vol)]                       # drop near-degenerate hull slivers only
    used = np.zeros(len(pts), bool); used[tets.ravel()] = True
    assert used.all()
    movable = pts[:, 2] > 1e-4                                         # bottom layer fixed ...
    U0 = np.zeros_like(pts)
    U0[~movable, 0] = 2e-6                                             # ... and sheared sideways
    ft = np.zeros_like(pts)
    top = pts[:, 2] > 9e-4
    ft[top, 2] = 5e-9
    mat = BATCH_A
    ref = c_oracle.relax(pts, tets, movable, ft, U0, beams, mat, 0.01, 40, 0.01, nthreads=1)
    with _lib.Context(pts, tets, movable, beams=beams) as ctx:
        info = ctx.info()
        ctx.set_material(*mat); ctx.set_targets(ft); ctx.set_displacements(U0)
        relrec, cgit, n = ctx.relax(0.01, 40, 0.01)
        U = ctx.displacements(); f = ctx.forces()
    assert info["cg_wmax"] > 18 and info["cg_resident"] == 1          # the shared-memory resident kernel, by necessity
    assert n[0] == ref["n_iter"]
    assert np.array_equal(U[~movable], U0[~movable])
    assert np.linalg.norm(U - ref["U"]) / np.linalg.norm(ref["U"]) < 1e-4
    assert np.abs(f - ref["forces"]).max() < 1e-3 * np.abs(ref["forces"]).max()
    # the sheared bottom layer holds tets with four fixed nodes and non-zero strain: E_glo must leave them out
    # on both sides (saenopy's rule, oracle `energy_rule`), so the logged energies agree from the first row on
    assert (~np.any(movable[tets], axis=1)).sum() > 0
    assert np.allclose(relrec[0][:, 1], ref["relrec"][:, 1], rtol=1e-6)


def test_hub_node_beyond_the_device_setup_capacity(beams, capfd, monkeypatch):
    """A node with more than 128 distinct neighbours (a hub at the centre of a point cloud on a sphere): the
    device-side adjacency declines and the host path builds the structure; results still follow the oracle."""
    from scipy.spatial import Delaunay
    from jointforces_b200 import _lib
    from oracle import c_oracle
    rng = np.random.default_rng(4)
    v = rng.normal(size=(260, 3))
    shell = v / np.linalg.norm(v, axis=1)[:, None] * 1e-3
    pts = np.vstack([np.zeros((1, 3)), shell, shell * 2.0])           # hub, inner shell, outer (fixed) shell
    tets = Delaunay(pts).simplices.astype(np.int32)
    B = pts[tets[:, 1:]] - pts[tets[:, :1]]
    vol = np.abs(np.linalg.det(B)) / 6
    tets = tets[vol > 1e-3 * np.median(vol)]                       # drop hull slivers
    assert len(np.unique(tets)) == len(pts)
    hub_degree = len(np.unique(tets[np.any(tets == 0, axis=1)]))
    assert hub_degree > 129                                          # more than ADJ_CAP neighbours incl. itself
    movable = np.ones(len(pts), bool)
    movable[261:] = False
    ft = np.zeros_like(pts)
    ft[1:261] = -shell / 1e-3 * 2e-7                                   # inner shell pulled towards the hub
    mat = COLLAGEN12
    ref = c_oracle.relax(pts, tets, movable, ft, np.zeros_like(pts), beams, mat, 0.0033, 15, 0.01, nthreads=1)
    monkeypatch.setenv("JF_TIMING", "1")
    with _lib.Context(pts, tets, movable, beams=beams) as ctx:
        info = ctx.info()
        ctx.set_material(*mat); ctx.set_targets(ft); ctx.set_displacements(np.zeros_like(pts))
        relrec, cgit, n = ctx.relax(0.0033, 15, 0.01)
        U = ctx.displacements()
    assert "adjacency (host)" in capfd.readouterr().err
    assert info["cg_wmax"] >= hub_degree - 1
    assert n[0] == ref["n_iter"]
    assert np.linalg.norm(U - ref["U"]) / np.linalg.norm(ref["U"]) < 1e-4


def test_golden_oracle_table_rows(beams):
    """GPU curves against the committed oracle curves of the BatchA recipe at 1 Pa .. 10 kPa (1e-3, BASELINE tolerance)."""
    from jointforces_b200.sweep import relax_load_cases
    from jointforces_b200.simulation import spherical_boundary_conditions, lookup_bins
    path = os.path.join(GOLDEN, "oracle_batchA_rows.npz")
    if not os.path.exists(path):
        pytest.skip("fixture missing")
    g = np.load(path)
    case = make_case(float(g["lf_iso"]), 1.0)
    r_in = 100 * 10 ** -6
    bcs = [spherical_boundary_conditions(case["coords"], r_in, 10000 * 10 ** -6, float(p)) for p in g["pressure"]]
    x, _ = lookup_bins()
    mat = dict(zip(("K_0", "D_0", "L_S", "D_S"), g["material"]))
    res, _ = relax_load_cases(case["coords"], case["tets"], bcs[0]["displacement"], [b["forces"] for b in bcs], mat,
                              float(g["step"]), 600, 0.01, beams=beams, curve_bins=(r_in, x), fetch_fields=False)
    for i, r in enumerate(res):
        assert len(r["relrec"]) - 1 == int(g["n_iter"][i])
        ok = ~np.isnan(g["curves"][i])
        assert np.array_equal(np.isnan(r["curve"]), ~ok)
        assert np.max(np.abs(r["curve"][ok] / g["curves"][i][ok] - 1)) < 1e-3, i


def test_single_reduction_kernel_orderings_agree_bitwise(beams, monkeypatch):
    """The register-resident single-reduction kernel in its three forms -- specialised instantiation with self-validating
    (tagged) Ap rows and no fence (default), the general instantiation (JF_CG_GENERAL=1), and fence-ordered rows
    (JF_CG_FENCED=1, with and without acquire polling) -- executes the same arithmetic: any row read before it arrived,
    or read torn, would change bits.  Repeated to give a race more than one chance."""
    from jointforces_b200 import _lib
    c = make_case(0.2, 300.0)                  # config-1 size: register-resident kernel, ~80 CTAs
    outs = {}
    for label, env in (("default", {}), ("general", {"JF_CG_GENERAL": "1"}), ("fenced", {"JF_CG_FENCED": "1"}),
                       ("fenced+acquire", {"JF_CG_FENCED": "1", "JF_CG_ACQUIRE": "1"}), ("default again", {})):
        for k in ("JF_CG_GENERAL", "JF_CG_FENCED", "JF_CG_ACQUIRE"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        with _lib.Context(c["coords"], c["tets"], c["movable"], beams=beams) as ctx:
            assert (ctx.info()["cg_resident"], ctx.info()["cg_fused"]) == (2, 1)
            ctx.set_material(*COLLAGEN12)
            ctx.set_targets(c["f_target"])
            ctx.set_displacements(np.zeros_like(c["coords"]))
            relrec, cgit, n = ctx.relax(0.0033, 40, 0.01)
            outs[label] = (ctx.displacements(), relrec[0], cgit[0])
    ref = outs["default"]
    for label, o in outs.items():
        assert np.array_equal(o[2], ref[2]), label
        assert np.array_equal(o[1], ref[1]), label
        assert np.array_equal(o[0], ref[0]), label


def test_on_chip_caches_of_the_streaming_kernels_are_transparent(beams, monkeypatch):
    """The streaming single-reduction kernel keeps the first block columns of each CTA's first group of slices in tensor
    memory and in shared memory (jf_cg_fused_stream_kernel<true>); the plain streaming kernel caches the first columns of one slice per warp.  A cached block is the same number read from another place and the summation order is unchanged, so every
    switch setting gives identical bits."""
    from jointforces_b200 import _lib
    c = make_case(0.2, 300.0)
    switches = ("JF_CG_STREAM_TMEM", "JF_CG_STREAM_NO_SMEM_CACHE", "JF_CG_CLASSIC", "JF_CG_NO_STREAM_HALO", "JF_CG_STREAM_NO_CACHE")

    def run(env, expect_fused):
        for k in switches:
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        with _lib.Context(c["coords"], c["tets"], c["movable"], beams=beams, flags=_lib.JF_FLAG_NO_RESIDENT) as ctx:
            assert (ctx.info()["cg_resident"], ctx.info()["cg_fused"]) == (0, expect_fused), ctx.info()
            ctx.set_material(*COLLAGEN12)
            ctx.set_targets(c["f_target"])
            ctx.set_displacements(np.zeros_like(c["coords"]))
            relrec, cgit, n = ctx.relax(0.0033, 12, 0.01)
            return ctx.displacements(), relrec[0], cgit[0]

    ref = run({}, 2)                                                   # tensor memory + shared memory (default)
    for env in ({"JF_CG_STREAM_TMEM": "0"}, {"JF_CG_STREAM_NO_SMEM_CACHE": "1"},
                {"JF_CG_STREAM_TMEM": "0", "JF_CG_STREAM_NO_SMEM_CACHE": "1"}):
        out = run(env, 2)
        assert np.array_equal(out[2], ref[2]) and np.array_equal(out[1], ref[1]) and np.array_equal(out[0], ref[0]), env
    plain = {"JF_CG_CLASSIC": "1", "JF_CG_NO_STREAM_HALO": "1"}        # the plain two-reduction streaming kernel
    ref = run(plain, 0)
    out = run(dict(plain, JF_CG_STREAM_NO_CACHE="1"), 0)
    assert np.array_equal(out[2], ref[2]) and np.array_equal(out[1], ref[1]) and np.array_equal(out[0], ref[0])
    for k in switches:
        monkeypatch.delenv(k, raising=False)
