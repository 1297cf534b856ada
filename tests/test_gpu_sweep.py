"""GPU tests of the jf.simulation call surface end to end: output-folder layout, the table
builder reading those folders, lockstep sweep vs oracle (BASELINE tolerance 1e-3 on the curves)."""
import os

import numpy as np
import pytest

import jointforces_b200 as jf
from conftest import COLLAGEN12

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def meshfile(tmp_path_factory):
    path = str(tmp_path_factory.mktemp("mesh") / "inclusion.msh")
    jf.mesh.spherical_inclusion(path, 100, 10000, 0.2)      # lf_iso 0.667: 360 nodes
    return path


def test_spherical_contraction_solver_writes_reference_files(meshfile, tmp_path, beams):
    from oracle import c_oracle
    out = str(tmp_path / "single")
    M = jf.simulation.spherical_contraction_solver(meshfile, out, 100.0, jf.materials.collagen12)
    assert sorted(os.listdir(out)) == ["parameters.txt", "relrec.txt", "solver.npz", "solver.saenopy"]
    par = jf.simulation.read_parameters(out)
    assert par["PRESSURE"] == 100.0 and abs(par["INNER_RADIUS"] - 100.0) < 1e-9 and par["TOTAL_NODES"] == 360
    relrec = np.loadtxt(out + "/relrec.txt")
    assert relrec.shape[1] == 3 and relrec[0, 0] == 0
    L = jf.simulation.load_solver(out + "/solver.saenopy")
    assert np.array_equal(L.mesh.displacements, M.mesh.displacements)
    # against the oracle on the same file contents
    coords, tets, r_in, r_out = jf.simulation.read_meshfile(meshfile)
    bc = jf.simulation.spherical_boundary_conditions(coords, r_in, r_out, 100.0)
    movable = np.any(np.isnan(bc["displacement"]), axis=1)
    ref = c_oracle.relax(coords, tets.astype(np.int32), movable, np.nan_to_num(bc["forces"]), np.zeros_like(coords),
                         beams, COLLAGEN12, 0.0033, 600, 0.01, nthreads=1)
    assert len(relrec) - 1 == ref["n_iter"]
    assert np.linalg.norm(L.mesh.displacements - ref["U"]) / np.linalg.norm(ref["U"]) < 1e-4
    assert np.abs(L.mesh.forces - ref["forces"]).max() < 1e-3 * np.abs(ref["forces"]).max()


def test_distribute_solver_layout_and_table(meshfile, tmp_path, beams):
    from oracle import c_oracle
    out = str(tmp_path / "sweep")
    const_args = {'meshfile': meshfile, 'outfolder': out, 'material': jf.materials.collagen12, 'max_iter': 600, 'step': 0.0033}
    seen = []
    jf.simulation.distribute_solver('jf.simulation.spherical_contraction_solver', const_args=const_args, var_arg='pressure',
                                    start=0.1, end=10000, n=7, log_scaling=True, n_cores=4, get_initial=True,
                                    callback=lambda i, n: seen.append((i, n)))
    assert 'outfolder' not in const_args                                  # the reference mutates the caller's dict
    values = np.loadtxt(out + "/pressure-values.txt")
    assert np.allclose(values, np.logspace(-1, 4, 7))
    assert sorted(os.listdir(out)) == ["pressure-values.txt"] + ["simulation%06d" % i for i in range(7)]
    assert sorted(seen) == [(i, 7) for i in range(7)]
    table = jf.simulation.create_lookup_table_solver(out)
    assert table['y'].shape == (7, 100) and np.allclose(table['pressure'], values)
    populated = ~np.isnan(table['y'][0])
    assert np.all(np.diff(table['y'][:, populated], axis=0) > 0)          # displacement grows with pressure at every bin
    # curve parity with the oracle at three pressures (1e-3 relative, BASELINE tolerance)
    coords, tets, r_in, r_out = jf.simulation.read_meshfile(meshfile)
    x, _ = jf.simulation.lookup_bins()
    for i in (0, 3, 6):
        bc = jf.simulation.spherical_boundary_conditions(coords, r_in, r_out, values[i])
        movable = np.any(np.isnan(bc["displacement"]), axis=1)
        ref = c_oracle.relax(coords, tets.astype(np.int32), movable, np.nan_to_num(bc["forces"]), np.zeros_like(coords),
                             beams, COLLAGEN12, 0.0033, 600, 0.01, nthreads=1)
        curve = jf.simulation.radial_median_curve(coords, ref["U"], r_in, x)
        ok = ~np.isnan(curve)
        assert np.max(np.abs(table['y'][i][ok] / curve[ok] - 1)) < 1e-3
    # the interpolants built from it behave
    get_displacement, get_pressure = jf.simulation.create_lookup_functions(table)
    xs = table['x'][populated][2]
    d = float(get_displacement(xs, 50.0))
    assert abs(float(get_pressure(xs, d)) / 50.0 - 1) < 0.02


def test_device_curves_equal_host_formula_bitwise(beams):
    """jf_get_curves vs the reference's numpy expression (simulation.py:296-299) on the same displacements."""
    from jointforces_b200 import _lib
    from jointforces_b200.mesh import generate_spherical_inclusion
    coords, tets = generate_spherical_inclusion(lf_iso=0.3)
    rng = np.random.default_rng(4)
    coords = coords * (1 + 0.2 * rng.random((len(coords), 1)))       # scatter the radii so that many bins are populated
    r_in = (100 * 10 ** -6 * 1e6) * 10 ** -6
    movable = np.linalg.norm(coords, axis=1) < 0.9e-2
    x, _ = jf.simulation.lookup_bins()
    with _lib.Context(coords, tets.astype(np.int32), movable, n_sys=2, beams=beams) as ctx:
        ctx.set_curve_bins(coords, r_in, x)
        Us = [rng.normal(size=coords.shape) * 1e-6, rng.normal(size=coords.shape) * 1e-5]
        Us[1][::7] = Us[1][1::7][: len(Us[1][::7])]                   # exact ties
        Us[0][::97] = 1e200                                           # |u|^2 overflows: infinite entries sort last, as in numpy
        for s, U in enumerate(Us):
            ctx.set_displacements(U, s)
        got = ctx.curves()
    for s, U in enumerate(Us):
        u = np.sqrt(np.sum(coords ** 2., axis=1)) / r_in
        v = np.sqrt(np.sum(U ** 2., axis=1)) / r_in
        with np.errstate(all="ignore"):
            expect = np.array([np.nanmedian(v[(u >= x[i]) & (u < x[i + 1])]) for i in range(len(x) - 1)])
        assert (~np.isnan(expect)).sum() > 40
        assert np.array_equal(np.isnan(got[s]), np.isnan(expect))
        assert np.array_equal(got[s][~np.isnan(expect)], expect[~np.isnan(expect)])


def test_opt_in_warm_start_chains_load_cases(beams):
    """warm_start=True: every load case starts from the previous one's relaxed field (kept on the device); the
    default stays the reference's effective behaviour, a cold start per pressure."""
    from jointforces_b200.sweep import relax_load_cases
    from jointforces_b200.mesh import generate_spherical_inclusion
    coords, tets = generate_spherical_inclusion(lf_iso=0.667)
    bcs = [jf.simulation.spherical_boundary_conditions(coords, 1e-4, 1e-2, p) for p in (50.0, 55.0, 60.0)]
    args = (coords, tets, bcs[0]["displacement"], [b["forces"] for b in bcs], jf.materials.collagen12, 0.0033, 600, 0.01)
    cold, _ = relax_load_cases(*args, beams=beams)
    warm, _ = relax_load_cases(*args, beams=beams, warm_start=True)
    n_cold = [len(r["relrec"]) - 1 for r in cold]
    n_warm = [len(r["relrec"]) - 1 for r in warm]
    assert n_warm[0] == n_cold[0] and np.array_equal(warm[0]["U"], cold[0]["U"])     # the first case has nothing to start from
    assert all(w < c / 2 for w, c in zip(n_warm[1:], n_cold[1:]))                     # later ones stop after a few steps
    for w, c in zip(warm[1:], cold[1:]):                                              # at nearly the same early-stopped field
        assert np.linalg.norm(w["U"] - c["U"]) < 0.1 * np.linalg.norm(c["U"])


# ---- against the reference's own functions (tests/golden/reference_simulation.npz) ---------------------------
@pytest.fixture(scope="module")
def ref_sim():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_simulation.npz"))


def test_device_curves_equal_reference_table_rows(ref_sim, tmp_path, beams):
    """The rows the reference's extract_deformation_curve_solver / create_lookup_table_solver computed from the
    golden displacement fields (simulation.py:263-323), reproduced by the device's radial median: same bits."""
    from jointforces_b200 import _lib
    (tmp_path / "mid.msh").write_text(str(ref_sim["sweep_msh"]))
    coords, tets, r_in, r_out = jf.simulation.read_meshfile(str(tmp_path / "mid.msh"))
    assert np.array_equal(coords, ref_sim["sweep_nodes"])
    bc = jf.simulation.spherical_boundary_conditions(coords, r_in, r_out, 1.0)
    movable = np.any(np.isnan(bc["displacement"]), axis=1)
    x, _ = jf.simulation.lookup_bins(1, 50, 100)
    with _lib.Context(coords, tets.astype(np.int32), movable, n_sys=7, beams=beams) as ctx:
        ctx.set_curve_bins(coords, (r_in * 1e6) * 10 ** -6, x)      # the radius as parameters.txt carries it (:296-297)
        for s in range(7):
            ctx.set_displacements(ref_sim["sweep_displacements"][s], s)
        got = ctx.curves()
    ref = ref_sim["table_y"]
    assert got.shape == ref.shape == (7, 100)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    assert np.array_equal(got[~np.isnan(ref)].view(np.uint64), ref[~np.isnan(ref)].view(np.uint64))


def test_sweep_files_equal_reference_files(ref_sim, tmp_path):
    """A real sweep on the golden mesh with the generator's arguments: pressure-values.txt and every parameters.txt
    are the texts the reference's distribute_solver / spherical_contraction_solver wrote."""
    (tmp_path / "mid.msh").write_text(str(ref_sim["sweep_msh"]))
    const_args = {"meshfile": str(tmp_path / "mid.msh"), "outfolder": str(tmp_path / "sweep"), "material": jf.materials.collagen12,
                  "max_iter": 77}
    jf.simulation.distribute_solver("ignored", const_args, var_arg="pressure", start=0.1, end=10000, n=7, log_scaling=True, n_cores=4)
    assert open(tmp_path / "sweep" / "pressure-values.txt").read() == str(ref_sim["sweep_values_txt"])
    assert sorted(os.listdir(tmp_path / "sweep")) == list(ref_sim["sweep_listing"])
    for i in range(7):
        folder = tmp_path / "sweep" / ("simulation%06d" % i)
        assert open(folder / "parameters.txt", encoding="utf-8").read() == str(ref_sim["sweep_parameters"][i])
        relrec = np.loadtxt(folder / "relrec.txt")
        assert relrec.ndim == 2 and relrec.shape[1] == 3 and 2 <= len(relrec) <= 78
    table = jf.simulation.create_lookup_table_solver(str(tmp_path / "sweep"))
    assert np.array_equal(table["pressure"], ref_sim["table_pressure"]) and np.array_equal(table["x"], ref_sim["table_x"])
    populated = ~np.isnan(ref_sim["table_y"][0])
    assert np.array_equal(~np.isnan(table["y"][0]), populated)           # same bins populated
    assert np.all(np.diff(table["y"][:, populated], axis=0) > 0)


def test_relrec_file_and_callback_follow_the_newton_steps(tmp_path, beams):
    """saenopy rewrites `relrecname` and calls `callback(solver, relrec, i, max_iterations)` after EVERY Newton step
    (what simulation.py:167 relies on for relrec.txt).  The device logs the rows while the host enqueues steps ahead;
    the caller's thread follows them: callbacks arrive for i = 0 .. n-1 in order with a growing record, and the file
    always holds the rows seen so far."""
    from conftest import make_case
    case = make_case(0.2, 100.0)                          # config-1 size: ~90 Newton steps, ~0.1 s
    M = jf.simulation.Solver()
    M.set_material_model(jf.simulation.SemiAffineFiberMaterial(*COLLAGEN12))
    M.set_nodes(case["coords"]); M.set_tetrahedra(case["tets"])
    M.set_boundary_condition(case["bc"]["displacement"], case["bc"]["forces"])
    name = str(tmp_path / "relrec.txt")
    calls, rows_in_file = [], []

    def cb(solver, relrec, i, max_iterations):
        calls.append((i, len(relrec), max_iterations))
        rows_in_file.append(len(np.atleast_2d(np.loadtxt(name))))
    rr = M.solve_boundarycondition(step_size=0.0033, max_iterations=600, rel_conv_crit=0.01, relrecname=name, callback=cb)
    n = len(rr) - 1
    assert [c[0] for c in calls] == list(range(n)) and all(c[2] == 600 for c in calls)
    assert [c[1] for c in calls] == list(range(2, n + 2))                      # start row + one row per finished step
    assert all(r >= c[1] for r, c in zip(rows_in_file, calls))                  # the file is never behind the callback
    assert len({r for r in rows_in_file}) > 3                                   # it really was rewritten while the solve ran
    assert np.array_equal(np.loadtxt(name), rr)
    # and the same solve without progress reporting gives the same bits
    M2 = jf.simulation.Solver()
    M2.set_material_model(jf.simulation.SemiAffineFiberMaterial(*COLLAGEN12))
    M2.set_nodes(case["coords"]); M2.set_tetrahedra(case["tets"])
    M2.set_boundary_condition(case["bc"]["displacement"], case["bc"]["forces"])
    rr2 = M2.solve_boundarycondition(step_size=0.0033, max_iterations=600, rel_conv_crit=0.01)
    assert np.array_equal(rr, rr2) and np.array_equal(M.mesh.displacements, M2.mesh.displacements)


def test_killed_solve_leaves_its_relaxation_record(tmp_path):
    """A solve killed half way (SIGKILL: no cleanup runs) has written relrec.txt for the steps it finished -- the
    reference gets this from saenopy's np.savetxt after every iteration."""
    import signal
    import subprocess
    import sys
    import time
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    name = str(tmp_path / "relrec.txt")
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import numpy as np, jointforces_b200 as jf\n"
            "from conftest import make_case, COLLAGEN12\n"
            "c = make_case(0.1, 100.0)\n"                  # config-4 size: a second or so of relaxation
            "M = jf.simulation.Solver(); M.set_material_model(jf.simulation.SemiAffineFiberMaterial(*COLLAGEN12))\n"
            "M.set_nodes(c['coords']); M.set_tetrahedra(c['tets']); M.set_boundary_condition(c['bc']['displacement'], c['bc']['forces'])\n"
            "print('go', flush=True)\n"
            "M.solve_boundarycondition(step_size=0.0033, max_iterations=600, rel_conv_crit=0.01, relrecname=%r)\n"
            "print('done', flush=True)\n" % (root, os.path.join(root, "tests"), name))
    proc = subprocess.Popen([sys.executable, "-c", code], stdout=subprocess.PIPE, text=True)
    assert proc.stdout.readline().strip() == "go"
    deadline = time.time() + 60
    while time.time() < deadline and not (os.path.exists(name) and os.path.getsize(name) > 300):
        time.sleep(0.01)
    proc.send_signal(signal.SIGKILL)
    proc.wait()
    assert proc.returncode == -signal.SIGKILL
    rows = np.atleast_2d(np.loadtxt(name))
    assert 2 <= len(rows) < 78 and rows.shape[1] == 3 and rows[0, 0] == 0 and np.all(rows[1:, 1] > 0)


def test_sweep_in_pairs_equals_sweep_one_case_at_a_time(tmp_path):
    """On the configs[2] mesh run_sweep relaxes two neighbouring pressures side by side (jf_cg_pair_kernel, the odd case
    alone).  Same Newton counts and, within the BASELINE curve tolerance (in fact at the 1e-5 trajectory-noise level),
    the same table rows and displacement fields as the sweep that relaxes one case at a time; folders complete."""
    from jointforces_b200 import sweep
    path = str(tmp_path / "cfg2.msh")
    jf.mesh.spherical_inclusion(path, 100, 10000, 0.05)
    values = np.array([0.3, 30.0, 40.0, 900.0, 1000.0])
    t2 = sweep.run_sweep(path, str(tmp_path / "pairs"), values, jf.materials.collagen12)
    t1 = sweep.run_sweep(path, str(tmp_path / "single"), values, jf.materials.collagen12, batch=1)
    assert np.array_equal(t2['index'], np.arange(5)) and np.allclose(t2['pressure'], values)
    ok = ~np.isnan(t1['y'])
    assert np.array_equal(np.isnan(t2['y']), ~ok)
    assert np.max(np.abs(t2['y'][ok] / t1['y'][ok] - 1)) < 1e-3
    for i in range(5):
        a, b = (np.loadtxt(str(tmp_path / d / ("simulation%06d" % i) / "relrec.txt")) for d in ("pairs", "single"))
        assert a.shape == b.shape                                           # same number of Newton steps
        Ua, Ub = (jf.simulation.load_solver(str(tmp_path / d / ("simulation%06d" % i) / "solver.saenopy")).mesh.displacements
                  for d in ("pairs", "single"))
        assert np.linalg.norm(Ua - Ub) / np.linalg.norm(Ub) < 1e-4
