"""Tier-2 parity: the oracle (and the CUDA path) against REAL saenopy -- runs only where ``import saenopy`` works.

saenopy (``/root/reference/setup.py:15``, ``>=0.7.4``) is absent from the build container and from the GPU boxes, so
every test here skips there; parity with real saenopy is UNPINNED until this file has run somewhere.
On a machine with saenopy:  ``python -m pytest tests/test_saenopy_ab.py -s``  prints ``saenopy.__version__`` and the
measured differences.  The solver is driven exactly as ``jointforces.simulation.spherical_contraction_solver`` drives
it (``simulation.py:100-104, 127-143, 167-168``).
"""
import os

import numpy as np
import pytest

from conftest import COLLAGEN12, LINEAR, make_case

saenopy = pytest.importorskip("saenopy")


def _saenopy_solve(case, mat, step, max_iter=600, conv_crit=0.01, relrecname=None):
    from saenopy.materials import SemiAffineFiberMaterial
    M = saenopy.Solver()                                                   # simulation.py:100
    M.set_material_model(SemiAffineFiberMaterial(*mat))                    # :101-102
    M.set_nodes(case["coords"])                                            # :103
    M.set_tetrahedra(case["tets"].astype(np.float64))                      # :104 (the reference passes floats, :71-76)
    M.set_boundary_condition(case["bc"]["displacement"], case["bc"]["forces"])   # :143
    rr = M.solve_boundarycondition(step_size=step, max_iterations=max_iter, rel_conv_crit=conv_crit, relrecname=relrecname)  # :167
    return M, np.asarray(rr if rr is not None else M.relrec)


def _report(name, **kw):
    print("\n[saenopy %s] %s: %s" % (saenopy.__version__, name, ", ".join("%s=%s" % kv for kv in kw.items())))


@pytest.mark.parametrize("lf_iso,mat,pressure,step", [(0.667, COLLAGEN12, 100.0, 0.0033), (0.667, LINEAR, 1.0, 0.066),
                                                      (0.2, COLLAGEN12, 100.0, 0.0033)])
def test_oracle_matches_saenopy(lf_iso, mat, pressure, step):
    """BASELINE tolerances between the CPU oracle and real saenopy: same Newton iteration count, displacements within
    1e-4 relative L2, energies of the relaxation record within 1e-4.  The oracle's closed-form material is tried first
    and the sampled-table emulation second; which one saenopy's answers sit closer to is reported."""
    from oracle import c_oracle
    from oracle.saeno_oracle import Solver, SemiAffineFiberMaterial, build_beams
    case = make_case(lf_iso, pressure)
    M, rr = _saenopy_solve(case, mat, step)
    U_ref = np.asarray(M.mesh.displacements)
    ref = c_oracle.relax(case["coords"], case["tets"], case["movable"], case["f_target"], np.zeros_like(case["coords"]),
                         build_beams(300), mat, step, 600, 0.01)
    err_closed = np.linalg.norm(ref["U"] - U_ref) / np.linalg.norm(U_ref)
    err_lut = None
    if lf_iso > 0.5:                                   # the numpy oracle with the sampled table: small mesh only
        S = Solver()
        S.set_material_model(SemiAffineFiberMaterial(*mat, mode="lut"))
        S.set_nodes(case["coords"]); S.set_tetrahedra(case["tets"])
        S.set_boundary_condition(case["bc"]["displacement"], case["bc"]["forces"])
        S.solve_boundarycondition(step_size=step, max_iterations=600, rel_conv_crit=0.01)
        err_lut = np.linalg.norm(S.mesh.displacements - U_ref) / np.linalg.norm(U_ref)
    _report("oracle vs saenopy, %d nodes" % len(case["coords"]), newton_saenopy=len(rr) - 1, newton_oracle=ref["n_iter"],
            err_closed_form="%.2e" % err_closed, err_table_emulation=("%.2e" % err_lut) if err_lut is not None else "n/a")
    assert ref["n_iter"] == len(rr) - 1
    assert min(err_closed, err_lut if err_lut is not None else 1.0) < 1e-4
    assert np.allclose(ref["relrec"][:, 1], rr[:, 1], rtol=1e-4)
    # E_glo: saenopy counts only tets with a free node (oracle `energy_rule`); first row logs sum f^2 of the start state
    assert abs(ref["relrec"][0, 2] - rr[0, 2]) <= 1e-6 * abs(rr[0, 2]) + 1e-300


def test_material_lookup_rule():
    """saenopy's sampled material table against the oracle's emulation of it (mode='lut', linear interpolation between
    samples): w, w', w'' at arguments on and between samples."""
    from saenopy.materials import SemiAffineFiberMaterial as Real
    from oracle.saeno_oracle import SemiAffineFiberMaterial
    lam = np.array([-0.0300004, -1.2345e-7, 0.0, 3.3e-7, 0.0040007, 0.0200003, 0.2500006])
    real = Real(*COLLAGEN12)
    lookup = real.generate_look_up_table() if hasattr(real, "generate_look_up_table") else real.lookUpEpsilon
    r0, r1, r2 = lookup(lam)
    best = {}
    for rule in ("interp", "ceil", "round", "floor"):
        o0, o1, o2 = SemiAffineFiberMaterial(*COLLAGEN12, mode="lut", lut_index=rule)(lam)
        best[rule] = max(np.max(np.abs(o1 - r1) / (np.abs(r1) + 1e-30)), np.max(np.abs(o2 - r2) / np.abs(r2)))
    _report("lookup rule", **{k: "%.1e" % v for k, v in best.items()})
    assert best["interp"] < 1e-9, best


def test_solver_files_round_trip(tmp_path):
    """Row f2 of SURVEY §8: a ``solver.saenopy`` written by this package loads in real saenopy (what the reference's
    table builder does, simulation.py:279-281), and one written by saenopy loads here."""
    from jointforces_b200 import solver as slv
    case = make_case(0.667, 100.0)
    M, rr = _saenopy_solve(case, COLLAGEN12, 0.0033, max_iter=3)
    M.save(str(tmp_path / "theirs.saenopy"))
    L = slv.load_solver(str(tmp_path / "theirs.saenopy"))
    assert np.array_equal(L.mesh.nodes, M.mesh.nodes) and np.array_equal(L.mesh.displacements, M.mesh.displacements)
    slv.save_solver_file(str(tmp_path / "ours.saenopy"), case["coords"], case["tets"].astype(np.int64), np.asarray(M.mesh.displacements),
                         np.asarray(M.mesh.forces), case["f_target"], case["movable"], ["SemiAffineFiberMaterial", *COLLAGEN12], rr)
    T = saenopy.load(str(tmp_path / "ours.saenopy")) if hasattr(saenopy, "load") else saenopy.solver.load_solver(str(tmp_path / "ours.saenopy"))
    assert np.array_equal(np.asarray(T.mesh.nodes), case["coords"])
    assert np.array_equal(np.asarray(T.mesh.displacements), np.asarray(M.mesh.displacements))
    _report("solver.saenopy round trip", theirs_keys=sorted(np.load(str(tmp_path / "theirs.saenopy"), allow_pickle=True).files)[:12])


def test_reference_table_builder_reads_our_folders(tmp_path):
    """The unmodified reference ``create_lookup_table_solver`` (where the reference package itself imports) on folders
    written by this package -- through ``solver.saenopy`` or its own ``solver.npz`` fallback (simulation.py:278-291)."""
    jointforces = pytest.importorskip("jointforces")
    from jointforces_b200 import solver as slv, simulation as sim
    case = make_case(0.667, 100.0)
    M, rr = _saenopy_solve(case, COLLAGEN12, 0.0033, max_iter=3)
    U = np.asarray(M.mesh.displacements)
    for i, p in enumerate((50.0, 100.0)):
        folder = tmp_path / ("simulation%06d" % i)
        os.makedirs(folder)
        bc = sim.spherical_boundary_conditions(case["coords"], case["r_in"], case["r_out"], p)
        sim.write_parameters(str(folder), dict(K_0=1645, D_0=0.0008, L_S=0.0075, D_S=0.033), p, bc, case["r_in"], case["r_out"], len(U))
        for name, layout in (("solver.saenopy", "saenopy"), ("solver.npz", "legacy")):
            slv.save_solver_file(str(folder / name), case["coords"], case["tets"].astype(np.int64), U * (i + 1), np.zeros_like(U),
                                 case["f_target"], case["movable"], ["SemiAffineFiberMaterial", *COLLAGEN12], rr, layout=layout)
    theirs = jointforces.simulation.create_lookup_table_solver(str(tmp_path))
    ours = sim.create_lookup_table_solver(str(tmp_path))
    assert np.array_equal(theirs["pressure"], ours["pressure"]) and np.array_equal(np.isnan(theirs["y"]), np.isnan(ours["y"]))
    assert np.array_equal(np.nan_to_num(theirs["y"]), np.nan_to_num(ours["y"]))


@pytest.mark.gpu
def test_cuda_path_matches_saenopy():
    """The product against real saenopy at the BASELINE tolerances (config-1 size)."""
    from jointforces_b200 import _lib
    from jointforces_b200.solver import build_beams
    case = make_case(0.2, 100.0)
    M, rr = _saenopy_solve(case, COLLAGEN12, 0.0033)
    out = _lib.solve_boundarycondition(case["coords"], case["tets"], case["movable"], case["f_target"], np.zeros_like(case["coords"]),
                                       COLLAGEN12, 0.0033, 600, 0.01, beams=build_beams(300))
    err = np.linalg.norm(out["U"] - np.asarray(M.mesh.displacements)) / np.linalg.norm(np.asarray(M.mesh.displacements))
    _report("CUDA vs saenopy", newton=(out["n_iter"], len(rr) - 1), err="%.2e" % err)
    assert out["n_iter"] == len(rr) - 1 and err < 1e-4
