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
    assert sorted(os.listdir(out)) == ["parameters.txt", "relrec.txt", "solver.saenopy"]
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
