"""Host side of the path against the reference's OWN functions.

tests/golden/reference_simulation.npz was produced by tests/golden/make_reference_functions.py, which compiles
spherical_contraction_solver, distribute_solver, extract_deformation_curve_solver, create_lookup_table_solver
(jointforces/simulation.py:81-323) and materials.py from their source files under /root/reference and runs them with
the saenopy object replaced by a recorder.  Everything here is bit for bit: arrays, file texts, call arguments.
"""
import os

import numpy as np
import pytest

import jointforces_b200 as jf
from jointforces_b200 import simulation as sim, sweep
from jointforces_b200.solver import save_solver_file

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "reference_simulation.npz")


@pytest.fixture(scope="module")
def g():
    return np.load(GOLDEN)


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b, equal_nan=True)


def test_materials_match_reference_module(g):
    for name in ("collagen06", "collagen12", "collagen24", "fibrin40", "matrigel100"):
        m = getattr(jf.materials, name)
        assert [m[k] for k in ("K_0", "D_0", "L_S", "D_S")] == list(g["material_" + name])
    assert [jf.materials.linear(394)[k] for k in ("K_0", "D_0", "L_S", "D_S")] == list(g["material_linear_394"])


class Recorder:
    last = None

    def __init__(self):
        self.calls = {}
        Recorder.last = self

    def set_material_model(self, m):
        self.calls["material"] = m

    def set_nodes(self, c):
        self.calls["nodes"] = np.array(c)

    def set_tetrahedra(self, t):
        self.calls["tets"] = np.array(t)

    def set_boundary_condition(self, d, f):
        self.calls["bc_displacement"], self.calls["bc_forces"] = np.array(d), np.array(f)

    def set_initial_displacements(self, u):
        self.calls["initial"] = np.array(u)

    def solve_boundarycondition(self, **kw):
        self.calls["solve"] = kw

    def save(self, path):
        self.calls["save"] = path


def test_spherical_contraction_solver_hands_over_what_the_reference_does(g, tmp_path, monkeypatch, capsys):
    """Same mesh file, same arguments: nodes, tets, NaN-coded boundary conditions, start values, solver keywords,
    file names and parameters.txt are the reference's, bit for bit."""
    (tmp_path / "small.msh").write_text(str(g["small_msh"]))
    monkeypatch.setattr(sim, "Solver", Recorder)
    monkeypatch.setattr(sim, "SemiAffineFiberMaterial", lambda *a: tuple(a))
    sim.spherical_contraction_solver(str(tmp_path / "small.msh"), str(tmp_path / "one"), 123.456, jf.materials.collagen12,
                                     None, None, False, g["one_U0"], 55, 0.01, 0.02)
    rec = Recorder.last.calls
    assert list(rec["material"]) == list(g["one_material"])
    assert same(rec["nodes"], g["one_nodes"]) and same(rec["tets"], g["one_tets"])
    assert same(rec["bc_displacement"], g["one_bc_displacement"])
    assert same(rec["bc_forces"], g["one_bc_forces"])
    assert same(rec["initial"], g["one_initial"])
    kwargs = repr(sorted((k, v if k != "relrecname" else os.path.basename(v)) for k, v in rec["solve"].items()))
    assert kwargs == str(g["one_solve_kwargs"])
    assert os.path.basename(rec["save"]) == str(g["one_save"])
    assert open(tmp_path / "one" / "parameters.txt", encoding="utf-8").read() == str(g["one_parameters"])
    out = capsys.readouterr().out
    assert "Inner node spacing: " in out and "Outer node spacing: " in out


def fake_engine(golden):
    """Stands in for the device (sweep.SweepEngine): returns the displacement fields the generator's recorder wrote,
    recognising the load case by its total inner-shell force."""
    forces_of = {}

    class Engine:
        batch = 1
        totals = None

        def __init__(self, coords, tets, bc_displacement, material, step, max_iter, conv_crit, device=0, batch=None, beams=None,
                     curve_bins=None, initial_displacements=None):
            assert same(coords, golden["sweep_nodes"]) and max_iter == 77 and (step, conv_crit) == (0.0033, 0.01)
            self.coords, (self.r_curve, self.x) = coords, curve_bins
            for i, v in enumerate(np.logspace(-1, 4, 7)):
                f = sim.spherical_boundary_conditions(coords, 100e-6, 10000e-6, float(v))["forces"]
                forces_of[float(np.nansum(np.abs(f)))] = i

        def relax_batch(self, forces_list, fetch_fields=True, keep_displacements=False):
            out = []
            for f in forces_list:
                key = float(np.nansum(np.abs(f)))
                i = forces_of[min(forces_of, key=lambda k: abs(k - key))]
                U = golden["sweep_displacements"][i]
                out.append({"U": U, "forces": np.zeros_like(U), "relrec": np.zeros((2, 3)), "cg_iters": np.zeros(1, int), "ms": 1.0,
                            "curve": sim.radial_median_curve(self.coords, U, self.r_curve, self.x)})
            return out

        def close(self):
            pass
    return Engine



def test_distribute_solver_files_and_table_match_reference(g, tmp_path, monkeypatch):
    """distribute_solver -> folders -> create_lookup_table_solver, with the device replaced by the golden fields:
    values file, folder names, every parameters.txt, callbacks and the table equal the reference's."""
    (tmp_path / "mid.msh").write_text(str(g["sweep_msh"]))
    monkeypatch.setattr(sweep, "SweepEngine", fake_engine(g))
    seen = []
    const_args = {"meshfile": str(tmp_path / "mid.msh"), "outfolder": str(tmp_path / "sweep"), "material": jf.materials.collagen12,
                  "max_iter": 77}
    sim.distribute_solver("ignored", const_args, var_arg="pressure", start=0.1, end=10000, n=7, log_scaling=True, n_cores=1,
                          callback=lambda i, n: seen.append((i, n)))
    assert repr(sorted(const_args)) == str(g["sweep_const_args_after"])
    assert open(tmp_path / "sweep" / "pressure-values.txt").read() == str(g["sweep_values_txt"])
    assert sorted(os.listdir(tmp_path / "sweep")) == list(g["sweep_listing"])
    # the reference calls back after every solve and once more at the end (:246-247) with (n, n)
    # the reference calls back (index, n) after every launch in index order (and once more with (n, n)); here the
    # solves finish most-expensive-first, so the same calls arrive in schedule order
    assert sorted(tuple(s) for s in seen) == sorted(tuple(s) for s in g["sweep_callbacks"])[:len(seen)] and len(seen) >= 7
    for i in range(7):
        folder = tmp_path / "sweep" / ("simulation%06d" % i)
        assert open(folder / "parameters.txt", encoding="utf-8").read() == str(g["sweep_parameters"][i])
        assert sorted(os.listdir(folder)) == ["parameters.txt", "relrec.txt", "solver.npz", "solver.saenopy"]
    table = sim.create_lookup_table_solver(str(tmp_path / "sweep"), 1, 50, 100)
    assert same(table["pressure"], g["table_pressure"]) and same(table["x"], g["table_x"]) and same(table["y"], g["table_y"])
    assert 10 < (~np.isnan(table["y"][0])).sum() < 100            # populated and empty bins both occur
    # linear spacing branch (the reference spaces the logarithms, :192)
    class Zero:
        batch, totals = 1, None

        def __init__(self, *a, **k):
            pass

        def relax_batch(self, forces_list, fetch_fields=True, keep_displacements=False):
            return [{"U": np.zeros((72, 3)), "forces": np.zeros((72, 3)), "relrec": np.zeros((1, 3)), "cg_iters": np.zeros(0, int),
                     "ms": 0.0, "curve": np.full(100, np.nan)} for _ in forces_list]

        def close(self):
            pass
    monkeypatch.setattr(sweep, "SweepEngine", Zero)
    (tmp_path / "small.msh").write_text(str(g["small_msh"]))
    sim.distribute_solver("ignored", {"meshfile": str(tmp_path / "small.msh"), "outfolder": str(tmp_path / "lin"),
                                      "material": jf.materials.collagen12}, start=1, end=100, n=3, log_scaling=False, n_cores=1)
    assert open(tmp_path / "lin" / "pressure-values.txt").read() == str(g["lin_values_txt"])


def test_curve_extraction_matches_reference(g, tmp_path):
    """extract_deformation_curve_solver on a folder holding the golden field: the reference's row, bit for bit
    (np.nanmedian over the same members; empty bins NaN)."""
    coords = g["sweep_nodes"]
    x, _ = sim.lookup_bins(1, 50, 100)
    for i in (0, 3, 6):
        folder = tmp_path / ("simulation%06d" % i)
        folder.mkdir()
        (folder / "parameters.txt").write_text(str(g["sweep_parameters"][i]), encoding="utf-8")
        U = g["sweep_displacements"][i]
        save_solver_file(str(folder / "solver.saenopy"), coords, np.zeros((1, 4), dtype=np.int64), U, np.zeros_like(U),
                         np.zeros_like(U), np.ones(len(coords), dtype=bool), {}, np.zeros((1, 3)))
        res = sim.extract_deformation_curve_solver(str(folder), x)
        assert res["pressure"] == g["table_pressure"][i]
        assert same(res["displacements"], g["table_y"][i])
