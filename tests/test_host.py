"""CPU tests of the host side: file formats and call surface kept from the reference, the
saenopy object protocol mirror, the C ABI's symbol table, sweep sharding and the table gather
(gloo, world_size 2).  No compute calls: this box has no GPU."""
import inspect
import os
import time
import re

import numpy as np
import pytest

import jointforces_b200 as jf
from jointforces_b200 import _lib, simulation as sim, solver as slv, sweep

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- materials (reference materials.py:1-41) ----------------------------------------------------
def test_materials_values():
    assert jf.materials.collagen06 == {'K_0': 447, 'D_0': 0.0008, 'L_S': 0.0075, 'D_S': 0.033}
    assert jf.materials.collagen12 == {'K_0': 1645, 'D_0': 0.0008, 'L_S': 0.0075, 'D_S': 0.033}
    assert jf.materials.collagen24 == {'K_0': 5208, 'D_0': 0.0008, 'L_S': 0.0075, 'D_S': 0.033}
    assert jf.materials.fibrin40 == {'K_0': 2091, 'D_0': 0.002, 'L_S': 1e30, 'D_S': 1e30}
    assert jf.materials.matrigel100 == {'K_0': 2364, 'D_0': 1e30, 'L_S': 1e30, 'D_S': 1e30}
    assert jf.materials.linear(394) == {'K_0': 394, 'D_0': 1e30, 'L_S': 1e30, 'D_S': 1e30}
    assert jf.materials.custom(1, 2, 3, 4) == {'K_0': 1, 'D_0': 2, 'L_S': 3, 'D_S': 4}


def test_call_surface_signatures():
    """Signatures verbatim; distribute_solver passes eleven arguments positionally."""
    p = list(inspect.signature(sim.spherical_contraction_solver).parameters)
    assert p == ['meshfile', 'outfolder', 'pressure', 'material', 'r_inner', 'r_outer', 'logfile',
                 'initial_displacenemts', 'max_iter', 'step', 'conv_crit']
    d = inspect.signature(sim.distribute_solver).parameters
    assert list(d) == ['func', 'const_args', 'var_arg', 'start', 'end', 'n', 'log_scaling', 'n_cores', 'get_initial',
                       'max_iter', 'step', 'conv_crit', 'callback']
    assert (d['start'].default, d['end'].default, d['n'].default) == (0.1, 1000, 120)
    assert (d['max_iter'].default, d['step'].default, d['conv_crit'].default) == (600, 0.0033, 0.01)
    assert list(inspect.signature(sim.create_lookup_table_solver).parameters) == ['folder', 'x0', 'x1', 'n']
    assert list(inspect.signature(sim.read_meshfile).parameters) == ['meshfile', 'r_inner', 'r_outer']
    for name in ("extract_deformation_curve_solver", "create_lookup_functions", "save_lookup_table",
                 "load_lookup_functions", "linear_lookup_interpolator", "Solver", "SemiAffineFiberMaterial", "load_solver"):
        assert hasattr(sim, name)


# ---- mesh file (reference simulation.py:25-78, mesh.py:86) -----------------------------------------
def test_meshfile_roundtrip(tmp_path):
    coords, tets = jf.mesh.spherical_inclusion(str(tmp_path / "m.msh"), 100, 10000, 0.2)
    text = open(tmp_path / "m.msh").read()
    assert text.endswith('$Jointforces\ninfo=This mesh was created with JOINTFORCES.\ntype=spherical_inclusion\n'
                         'r_inner=100\nr_outer=10000\nlength_factor=0.2\n$EndJointforces\n')
    c, t, ri, ro = sim.read_meshfile(str(tmp_path / "m.msh"), 5, 7)     # radii in the file win
    assert ri == 100 * 10 ** -6 and ro == 10000 * 10 ** -6                # the reference's own expression (simulation.py:54-55)
    assert t.dtype == np.float64 and t.min() == 0 and t.max() == len(coords) - 1   # float, 0-based, like the reference
    assert np.array_equal(c, coords) and np.array_equal(t.astype(int), tets)
    assert len(np.unique(t)) == len(coords)                                          # no orphan nodes


@pytest.mark.parametrize("name", ["gmsh_like", "with_block", "ragged", "generated"])
def test_meshfile_matches_reference_function(tmp_path, name):
    """The reference's own read_meshfile (simulation.py:25-78), run by tests/golden/make_reference_functions.py:
    both readers (the library's threaded one and the line loop it falls back to) return its arrays bit for bit."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_meshfile.npz"))
    path = tmp_path / (name + ".msh")
    path.write_text(str(g[name + "_text"]))
    fast = sim._read_meshfile_fast(str(path), 7.0, 9.0)
    assert fast is not None, "the threaded reader must take plain decimal files"
    for c, t, ri, ro in (fast, sim._read_meshfile_lines(str(path), 7.0, 9.0), sim.read_meshfile(str(path), 7.0, 9.0)):
        assert c.dtype == np.float64 and t.dtype == np.float64
        assert np.array_equal(c.view(np.uint64), g[name + "_coords"].view(np.uint64))     # same bits, signed zeros included
        assert np.array_equal(t, g[name + "_tets"])
        assert (ri, ro) == tuple(g[name + "_radii"])


def test_meshfile_fast_reader_leaves_irregular_files_to_the_line_loop(tmp_path):
    """Whatever is not a plain decimal ASCII file takes the reference's loop, with the reference's outcome."""
    base = "$Nodes\n4\n1 0 0 0\n2 1 0 0\n3 0 1 0\n4 0 0 1\n$EndNodes\n$Elements\n1\n1 4 0 1 2 3 4\n$EndElements\n"
    variants = {
        "crlf": base.replace("\n", "\r\n"),                       # universal newlines: the reference parses it
        "float_ids": base.replace("1 4 0 1 2 3 4", "1 4 0 1.0 2.0 3.0 4.0"),
        "underscore": base.replace("2 1 0 0", "2 1_0 0 0"),         # float('1_0') == 10.0 in Python
        "nan_coord": base.replace("2 1 0 0", "2 nan 0 0"),
        "unicode": base + "$Comment\n\u00b5m\n$EndComment\n",
    }
    for name, text in variants.items():
        path = tmp_path / (name + ".msh")
        path.write_text(text, newline="") if name == "crlf" else path.write_text(text, encoding="utf-8")
        assert sim._read_meshfile_fast(str(path), 1, 2) is None, name
        c, t, ri, ro = sim.read_meshfile(str(path), 1, 2)
        assert c.shape == (4, 3) and np.array_equal(t, [[0, 1, 2, 3]]), name
    assert sim.read_meshfile(str(tmp_path / "underscore.msh"), 1, 2)[0][1, 0] == 10.0
    # errors are the line loop's errors
    for name, text, exc in (("short_node", base.replace("2 1 0 0", "2 1 0"), ValueError),
                            ("no_nodes", base.replace("$Nodes", "$Knots"), ValueError),
                            ("too_few", base.replace("$Elements\n1", "$Elements\n5"), ValueError),
                            ("truncated", base.replace("$Elements\n1", "$Elements\n5").replace("$EndElements\n", ""), IndexError),
                            ("block_without_radii", base + "$Jointforces\ninfo=x\n$EndJointforces\n", KeyError)):
        path = tmp_path / (name + ".msh")
        path.write_text(text)
        assert sim._read_meshfile_fast(str(path), 1, 2) is None, name
        with pytest.raises(exc):
            sim.read_meshfile(str(path), 1, 2)
    with pytest.raises(FileNotFoundError):
        sim.read_meshfile(str(tmp_path / "missing.msh"), 1, 2)


def test_meshfile_fast_reader_large_file_equals_line_loop(tmp_path):
    """60 k tets with full-precision coordinates: identical arrays, int32 connectivity on request."""
    from jointforces_b200 import _lib
    coords, tets = jf.mesh.spherical_inclusion(str(tmp_path / "m.msh"), 100, 10000, 0.06)
    a = sim._read_meshfile_fast(str(tmp_path / "m.msh"), None, None)
    b = sim._read_meshfile_lines(str(tmp_path / "m.msh"))
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2:] == b[2:]
    assert np.array_equal(a[0], coords) and len(a[1]) == len(tets) > 50000
    for threads in (1, 3):
        c32, t32, radii = _lib.read_msh(str(tmp_path / "m.msh"), want_int32=True, n_threads=threads)
        assert t32.dtype == np.int32 and np.array_equal(t32, tets) and np.array_equal(c32, coords) and radii == (100.0, 10000.0)


def test_meshfile_without_trailer_and_ragged_elements(tmp_path):
    path = tmp_path / "r.msh"
    path.write_text("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n4\n1 0 0 0\n2 1 0 0\n3 0 1 0\n4 0 0 1\n$EndNodes\n"
                    "$Elements\n2\n1 4 2 4 3 1 2 3 4\n2 4 3 4 3 9 4 3 2 1\n$EndElements\n")
    with pytest.raises(ValueError, match="r_inner not defined"):
        sim.read_meshfile(str(path))
    with pytest.raises(ValueError, match="r_outer not defined"):
        sim.read_meshfile(str(path), r_inner=1)
    c, t, ri, ro = sim.read_meshfile(str(path), 1, 2)
    assert ri == 1 * 10 ** -6 and ro == 2 * 10 ** -6
    assert np.array_equal(t, [[0, 1, 2, 3], [3, 2, 1, 0]])                            # last four tokens of every line


def test_icoshell_mesh_is_conforming():
    pts, tets = jf.mesh.generate_spherical_inclusion(lf_iso=0.4, method="icoshell")
    faces = np.sort(np.concatenate([tets[:, [0, 1, 2]], tets[:, [0, 1, 3]], tets[:, [0, 2, 3]], tets[:, [1, 2, 3]]]), axis=1)
    _, counts = np.unique(faces, axis=0, return_counts=True)
    assert set(counts) == {1, 2}                                      # interior faces shared by exactly two tets
    B = pts[tets[:, 1:]] - pts[tets[:, :1]]
    vol = np.abs(np.linalg.det(B)).sum() / 6
    assert abs(vol / (4 / 3 * np.pi * (1e-2 ** 3 - 1e-4 ** 3)) - 1) < 0.08


# ---- boundary conditions and parameters.txt (reference simulation.py:107-164, 265-275) ---------------
def test_boundary_conditions_and_parameter_file(tmp_path):
    coords, tets = jf.mesh.generate_spherical_inclusion(lf_iso=0.667)
    bc = sim.spherical_boundary_conditions(coords, 100e-6, 10000e-6, 250.0)
    inner, outer = bc['mask_inner'], bc['mask_outer']
    assert inner.sum() == 30 and outer.sum() == 30
    assert np.isnan(bc['displacement'][~outer]).all() and np.all(bc['displacement'][outer] == 0)
    assert np.isnan(bc['forces'][outer]).all() and np.all(bc['forces'][~outer & ~inner] == 0)
    fpn = 250.0 * 4 * np.pi * (100e-6) ** 2 / 30
    assert np.allclose(np.linalg.norm(bc['forces'][inner], axis=1), fpn, rtol=1e-13)
    assert np.allclose(bc['forces'][inner] / fpn, coords[inner] / 100e-6, rtol=1e-9)   # along +r (contraction)
    sim.write_parameters(str(tmp_path), jf.materials.collagen12, 250.0, bc, 100e-6, 10000e-6, len(coords))
    lines = open(tmp_path / "parameters.txt", encoding="utf-8").read().split("\n")
    assert len(lines) == 12
    assert lines[0] == "K_0 = 1645" and lines[4] == "PRESSURE = 250.0"
    assert lines[6] == "INNER_RADIUS = 100.0 µm" and lines[10] == "SURFACE_NODES = 30" and lines[11] == "TOTAL_NODES = 360"
    par = sim.read_parameters(str(tmp_path))
    assert par['PRESSURE'] == 250.0 and par['INNER_RADIUS'] == 100.0 and par['K_0'] == 1645


# ---- solver protocol ----------------------------------------------------------------------
def test_solver_protocol_host_side(tmp_path):
    M = slv.Solver()
    M.set_material_model(slv.SemiAffineFiberMaterial(1645, 0.0008, 0.0075, 0.033))
    with pytest.raises(ValueError):
        M.set_nodes(np.zeros((5, 2)))
    nodes = np.array([[0., 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 1]])
    M.set_nodes(nodes)
    with pytest.raises(ValueError):
        M.set_tetrahedra(np.array([[0., 1, 2, 9]]))
    M.set_tetrahedra(np.array([[0., 1, 2, 3], [1., 2, 3, 4]]))          # floats, as the reference passes them
    assert M.mesh.tetrahedra.dtype == np.int64
    disp = np.full((5, 3), np.nan)
    disp[4] = [0, 0, 0.5]
    forces = np.zeros((5, 3))
    forces[4] = np.nan
    forces[0] = [1, 2, 3]
    M.set_boundary_condition(disp, forces)
    assert M.mesh.movable.tolist() == [True, True, True, True, False]
    assert np.array_equal(M.mesh.displacements[4], [0, 0, 0.5]) and np.all(M.mesh.displacements[:4] == 0)
    assert np.array_equal(M.mesh.forces_target[0], [1, 2, 3]) and np.all(M.mesh.forces_target[4] == 0)
    U0 = np.ones((5, 3))
    M.set_initial_displacements(U0)
    assert np.all(M.mesh.displacements[:4] == 1) and np.array_equal(M.mesh.displacements[4], [0, 0, 0.5])
    bad = forces.copy(); bad[1, 0] = np.nan
    with pytest.raises(ValueError):
        M.set_boundary_condition(disp, bad)
    # save / load (simulation.py:168, 279-281) and legacy attribute names (:213, 286-287)
    M.mesh.forces = np.arange(15.).reshape(5, 3)
    M.relrec = np.array([[0, 1, 2.]])
    M.save(str(tmp_path / "solver.saenopy"))
    assert os.path.exists(tmp_path / "solver.saenopy") and not os.path.exists(str(tmp_path / "solver.saenopy") + ".npz")
    L = slv.load_solver(str(tmp_path / "solver.saenopy"))
    assert np.array_equal(L.mesh.nodes, nodes) and np.array_equal(L.mesh.displacements, M.mesh.displacements)
    assert np.array_equal(L.R, nodes) and np.array_equal(L.U, M.mesh.displacements) and np.array_equal(L.f, M.mesh.forces)
    assert L.material_model.as_tuple() == (1645.0, 0.0008, 0.0075, 0.033)
    # saenopy's Saveable layout as recalled: group/field keys, the material as a flattened list, and no key its
    # constructors would not take (extra keys make Solver(**data) raise)
    z = np.load(str(tmp_path / "solver.saenopy"))
    assert set(z.files) == {"mesh/nodes", "mesh/tetrahedra", "mesh/displacements", "mesh/forces", "mesh/forces_target", "mesh/movable",
                            "material_parameters", "material_parameters/0", "material_parameters/1", "material_parameters/2",
                            "material_parameters/3", "material_parameters/4", "relrec"}
    assert str(z["material_parameters"]) == "list" and str(z["material_parameters/0"]) == "SemiAffineFiberMaterial"
    assert [float(z["material_parameters/%d" % i]) for i in range(1, 5)] == [1645.0, 0.0008, 0.0075, 0.033]
    assert M.material_parameters == ["SemiAffineFiberMaterial", 1645.0, 0.0008, 0.0075, 0.033] == L.material_parameters
    # the legacy archive the reference's last loader fallback reads (simulation.py:252-258): np.load keys R, U, f
    M.save(str(tmp_path / "solver.npz"), layout="legacy")
    z = np.load(str(tmp_path / "solver.npz"), allow_pickle=True)
    assert np.array_equal(z["R"], nodes) and np.array_equal(z["U"], M.mesh.displacements) and np.array_equal(z["f"], M.mesh.forces)
    L2 = slv.load_solver(str(tmp_path / "solver.npz"))
    assert np.array_equal(L2.R, nodes) and np.array_equal(L2.mesh.displacements, M.mesh.displacements)
    with pytest.raises(FileNotFoundError):
        slv.load_solver(str(tmp_path / "nope.saenopy"))


def test_beams_equal_oracle():
    from oracle.saeno_oracle import build_beams
    assert np.array_equal(slv.build_beams(300), build_beams(300))


# ---- lookup table (reference simulation.py:263-323, 595-621) ---------------------------------------------
def test_radial_median_curve_equals_reference_formula():
    rng = np.random.default_rng(0)
    coords = rng.normal(size=(4000, 3)) * 1e-3
    displ = rng.normal(size=(4000, 3)) * 1e-6
    x, xc = sim.lookup_bins()
    u = np.sqrt(np.sum(coords ** 2., axis=1)) / 100e-6
    v = np.sqrt(np.sum(displ ** 2., axis=1)) / 100e-6
    with np.errstate(all="ignore"):
        expect = np.array([np.nanmedian(v[(u >= x[i]) & (u < x[i + 1])]) for i in range(len(x) - 1)])   # simulation.py:299
    got = sim.radial_median_curve(coords, displ, 100e-6, x)
    assert np.array_equal(np.isnan(got), np.isnan(expect)) and np.allclose(got[~np.isnan(got)], expect[~np.isnan(expect)], rtol=0, atol=0)


def _fake_sweep(folder, pressures):
    coords, tets = jf.mesh.generate_spherical_inclusion(lf_iso=0.4)
    r = np.linalg.norm(coords, axis=1)
    for i, p in enumerate(pressures):
        sub = os.path.join(folder, "simulation" + str(i).zfill(6))
        os.makedirs(sub)
        bc = sim.spherical_boundary_conditions(coords, 100e-6, 10000e-6, p)
        sim.write_parameters(sub, jf.materials.collagen12, p, bc, 100e-6, 10000e-6, len(coords))
        U = -coords / r[:, None] * (1e-6 * p ** 0.7 * (100e-6 / r) ** 2)[:, None]
        slv.save_solver_file(sub + "/solver.saenopy", coords, tets, U, np.zeros_like(U), np.zeros_like(U),
                             r < 0.999e-2, None, None)
    return coords


def test_lookup_table_from_folders_and_interpolants(tmp_path):
    pressures = np.logspace(-1, 3, 9)
    _fake_sweep(str(tmp_path), pressures)
    np.savetxt(str(tmp_path / "pressure-values.txt"), pressures)         # a file next to the folders must not disturb the glob
    table = sim.create_lookup_table_solver(str(tmp_path))
    assert table['y'].shape == (9, 100) and table['x'].shape == (100,)
    assert np.allclose(table['pressure'], pressures)
    ok = ~np.isnan(table['y'][0])
    assert ok.sum() >= 10      # nodes of the synthetic mesh sit on discrete shells: most 4 %-wide bins are empty (NaN, masked later)
    expect = 1e-6 * pressures[3] ** 0.7 / 100e-6 / table['x'][ok] ** 2
    assert np.allclose(table['y'][3][ok], expect, rtol=0.1)               # median over a 4 %-wide bin
    sim.save_lookup_table(table, str(tmp_path / "t.npy"))
    get_displacement, get_pressure = sim.load_lookup_functions(str(tmp_path / "t.npy"))
    d = float(get_displacement(2.0, 10.0))
    assert abs(float(get_pressure(2.0, d)) / 10.0 - 1) < 0.05
    assert jf.load(str(tmp_path / "t.npy"))['y'].shape == (9, 100)


def test_linear_lookup_interpolator_rescales(tmp_path):
    ref = {'pressure': np.array([1., 2.]), 'x': np.array([1., 2., 3.]), 'y': np.arange(6.).reshape(2, 3)}
    np.save(str(tmp_path / "linear_reference_k2364.npy"), ref)
    out = sim.linear_lookup_interpolator(250, str(tmp_path / "new.npy"), reference_folder=str(tmp_path))
    assert np.allclose(out['y'], ref['y'] * 2364 / (250 * 6))                   # K_0 = 6 E (simulation.py:707,733)
    assert np.allclose(jf.load(str(tmp_path / "new.npy"))['y'], out['y'])


# ---- sweep sharding and the gather ------------------------------------------------------------------------
def test_shard_indices_partition_and_balance():
    """Static schedule: longest-processing-time-first on the cost model -- a partition, and balanced in COST (not count)."""
    values = np.logspace(-1, 4, 150)
    cost = sweep.pressure_cost(values, {"K_0": 1645.0})
    assert cost.min() >= 1.0 and 1.3 < cost.max() < 3.0 and cost[0] < cost.max() > cost[-1]    # dearest in the middle decades
    for world in (1, 2, 4, 8):
        parts = [sweep.shard_indices(values, r, world) for r in range(world)]
        assert sorted(np.concatenate(parts).tolist()) == list(range(150))
        loads = np.array([cost[p].sum() for p in parts])
        assert loads.max() - loads.min() <= cost.max()                      # LPT bound: within one item
        assert loads.max() / loads.mean() < 1.02
    # a stiffer material shifts the curve along the pressure axis
    assert np.argmax(sweep.pressure_cost(values, {"K_0": 5208.0})) > np.argmax(cost)


def _queue_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    taken = []
    for sweep_no in range(2):                                   # two sweeps: each gets its own counter
        values = np.logspace(-1, 4, 23)
        cost = sweep.pressure_cost(values)
        order = np.argsort(-cost, kind="stable")
        claimer = sweep._Claimer(order, cost, rank, world, "dynamic")
        assert claimer.store is not None
        mine = []
        while True:
            chunk = claimer.take(2 if rank else 1)              # ranks may draw different chunk sizes
            if not chunk:
                break
            mine += chunk
            time.sleep(0.002 * (1 + rank))                      # uneven speeds: the faster rank ends up with more
        taken.append(mine)
        dist.barrier()
    np.savez(os.path.join(out_dir, "q%d.npz" % rank), a=np.array(taken[0]), b=np.array(taken[1]))
    dist.destroy_process_group()


def test_dynamic_work_queue_two_ranks_gloo(tmp_path):
    """The shared counter in the process group's store hands every load case to exactly one rank, most expensive first."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_queue_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    q = [np.load(tmp_path / ("q%d.npz" % r)) for r in range(2)]
    for key in ("a", "b"):
        both = np.concatenate([q[0][key], q[1][key]])
        assert sorted(both.tolist()) == list(range(23))
        assert len(q[0][key]) > 0 and len(q[1][key]) > 0
    cost = sweep.pressure_cost(np.logspace(-1, 4, 23))
    first = [int(q[r]["a"][0]) for r in range(2)]
    assert set(first) <= set(np.argsort(-cost, kind="stable")[:3].tolist())   # the dearest cases go out first


def test_work_order_keeps_neighbouring_pressures_together():
    """Batches relaxed side by side are runs of neighbours on the pressure axis (similar Newton / CG iteration counts), the
    dearest batch first; a batch size of one is plain cost order; every case appears exactly once; the static deal keeps
    batches on one rank."""
    values = np.logspace(-1, 4, 150)
    cost = sweep.pressure_cost(values)
    assert np.array_equal(sweep.work_order(values, cost, 1), np.argsort(-cost, kind="stable"))
    order = sweep.work_order(values, cost, 2)
    assert sorted(order.tolist()) == list(range(150))
    pairs = order.reshape(-1, 2)
    assert np.all(np.abs(pairs[:, 0] - pairs[:, 1]) == 1)                 # neighbours in the (sorted) value list
    pair_cost = cost[pairs].sum(1)
    assert np.all(np.diff(pair_cost) <= 1e-12)                            # dearest first
    shuffled = np.random.default_rng(1).permutation(150)
    o2 = sweep.work_order(values[shuffled], cost[shuffled], 2).reshape(-1, 2)
    assert np.all(np.abs(np.log10(values[shuffled][o2[:, 0]] / values[shuffled][o2[:, 1]])) < 0.04)   # unsorted input: still neighbours
    odd = sweep.work_order(values[:7], cost[:7], 2)
    assert sorted(odd.tolist()) == list(range(7)) and len(odd) == 7
    for rank in range(3):                                                 # static deal (no process group): whole batches per rank
        c = sweep._Claimer(order, cost, rank, 3, "lpt", batch=2)
        mine = np.asarray(c.static).reshape(-1, 2)
        assert np.all(np.abs(mine[:, 0] - mine[:, 1]) == 1)
    dealt = sorted(i for rank in range(3) for i in sweep._Claimer(order, cost, rank, 3, "lpt", batch=2).static)
    assert dealt == list(range(150))


def _gloo_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    values = np.logspace(-1, 4, 11)
    mine = sweep.shard_indices(values, rank, world)
    curves = np.outer(values[mine], np.arange(1, 101.0))
    curves[:, 7] = np.nan                                         # empty bins travel as NaN
    table = sweep.gather_table(mine, values[mine], curves, len(values), 100)
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), **table)
    dist.destroy_process_group()


def test_gather_table_two_ranks_gloo(tmp_path):
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    values = np.logspace(-1, 4, 11)
    for r in range(2):
        t = np.load(tmp_path / ("r%d.npz" % r))
        assert np.array_equal(t['index'], np.arange(11)) and np.allclose(t['pressure'], values)
        expect = np.outer(values, np.arange(1, 101.0)); expect[:, 7] = np.nan
        assert np.array_equal(np.isnan(t['y']), np.isnan(expect))
        assert np.allclose(np.nan_to_num(t['y']), np.nan_to_num(expect))


def test_gather_table_without_process_group_is_local():
    t = sweep.gather_table([2, 5], [1.0, 2.0], np.ones((2, 100)), 10, 100)
    assert t['index'].tolist() == [2, 5] and t['y'].shape == (2, 100)


# ---- C ABI ----------------------------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "jf_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(jf_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS)
    L = _lib.load()
    for name in declared:
        assert hasattr(L, name), name
    assert L.jf_version() >= 100


def test_no_cpu_fallback_without_device():
    """On a box without a GPU every compute entry must fail loudly (JF_ERR_CUDA), never fall back."""
    try:
        n = _lib.device_count()
    except _lib.JFError:
        n = 0
    if n > 0:
        pytest.skip("a CUDA device is present")
    nodes = np.array([[0., 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]])
    with pytest.raises(_lib.JFError, match="CUDA"):
        _lib.Context(nodes, np.array([[0, 1, 2, 3]], np.int32), np.ones(4, np.uint8))
    with pytest.raises(_lib.JFError):
        _lib.solve_boundarycondition(nodes, np.array([[0, 1, 2, 3]], np.int32), np.ones(4, np.uint8), np.zeros((4, 3)),
                                     np.zeros((4, 3)), (1.0, 1.0, 1.0, 1.0), 0.1, 1, 0.01)
    M = slv.Solver()
    M.set_material_model(slv.SemiAffineFiberMaterial(1, 1, 1, 1))
    M.set_nodes(nodes); M.set_tetrahedra(np.array([[0, 1, 2, 3]]))
    with pytest.raises(_lib.JFError):
        M.solve_boundarycondition()
    # the table interpolants and everything behind them live on the device as well
    table = {'pressure': np.array([1.0, 10.0, 100.0]), 'x': np.array([1.0, 2.0, 4.0]),
             'y': np.array([[1e-3, 5e-4, 2e-4], [1e-2, 5e-3, 2e-3], [1e-1, 5e-2, 2e-2]])}
    with pytest.raises(_lib.JFError, match="CUDA"):
        jf.force.create_lookup_functions(table)


def test_force_host_side_logic():
    """File ordering, argument checks and the callable protocol of the force module (no device needed)."""
    from jointforces_b200 import force
    names = ["def10.npy", "def2.npy", "def000001.npy", "defA3.npy"]
    assert sorted(names, key=force._natural_key) == ["def000001.npy", "def2.npy", "def10.npy", "defA3.npy"]
    assert force._filter_args(None) == (0, 0.0)
    assert force._filter_args(20) == (1, float(np.cos(2 * np.pi * 20 / 360)))
    with pytest.raises(TypeError, match="device"):
        force.infer_pressure([1.0], [1.0], [1.0], [1.0], 0.0, 0.0, 1.0, lambda a, b: a)
    with pytest.raises(TypeError):
        force.reconstruct_frames([], lambda a, b: a)
    assert len(force.COLUMNS) == 12 and force.SECTOR_ANGLES[0] == -180 and force.SECTOR_ANGLES[-1] == 175 and len(force.SECTOR_ANGLES) == 72
    import inspect
    assert list(inspect.signature(force.infer_pressure).parameters) == ['x_rav', 'y_rav', 'u_rav', 'v_rav', 'x_sph', 'y_sph', 'r_sph',
                                                                        'get_pressure', 'angle_filter']
    assert list(inspect.signature(force.reconstruct).parameters)[:8] == ['folder', 'lookupfile', 'muperpixel', 'outfile', 'r_min',
                                                                         'angle_filter', 'r_max', 'continuous_radii']


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "jointforces_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert "oracle" not in text.replace("# noqa", ""), "%s mentions the oracle" % f
