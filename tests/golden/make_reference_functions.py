"""Golden vectors from the reference's OWN Python functions, executed here.

`import jointforces` fails in this container (saenopy, matplotlib, ... are absent), but the functions on either
side of the solve only need numpy/scipy.  This script compiles them from their source files where they lie under
/root/reference (ast, nothing is copied into the repo), runs them on small inputs and stores inputs + outputs:

    python tests/golden/make_reference_functions.py

  read_meshfile            jointforces/simulation.py:25-78    -> reference_meshfile.npz (+ the .msh texts inside)
  create_lookup_functions  jointforces/simulation.py:595-621  -> reference_lookup.npz
  infer_pressure           jointforces/force.py:399-432       -> reference_lookup.npz

/root/reference does not exist on the GPU box: the tests read only the .npz files.
"""
import ast
import contextlib
import io
import os
import sys

import numpy as np

REF = "/root/reference/jointforces"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def reference_function(relpath, name, extra_globals=None):
    """The function `name` of a reference source file, compiled from the file in place."""
    path = os.path.join(REF, relpath)
    tree = ast.parse(open(path).read(), filename=path)
    node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)
    g = {"np": np}
    g.update(extra_globals or {})
    exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), g)
    return g[name]


MESH_TEXTS = {
    # gmsh-like: header sections, tags on the element lines, exponent notation, tabs, blank padding
    "gmsh_like": "$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n5\n1 0 0 0\n2 1.5e-1 0 0\n3 0 -2.5E+0 0\n4 0\t0 3.0000000000000004\n"
                 "5   0.1 0.2   0.30000000000000004\n$EndNodes\n$Elements\n2\n1 4 2 1 1 1 2 3 4\n2 4 2 1 1 2 3 4 5\n$EndElements\n",
    # with the trailer mesh.py writes; radii in the file win over the arguments
    "with_block": "$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n4\n1 1 0 0\n2 0 1 0\n3 0 0 1\n4 0.25 0.25 0.25\n$EndNodes\n$Elements\n1\n"
                  "1 4 0 1 2 3 4\n$EndElements\n$Jointforces\nr_inner=100\nr_outer=20000.5\nnote=hello\n$EndJointforces\n",
    # no trailing newline on the last element line, "+" signs, more than four node tokens ahead of the last four
    "ragged": "$Nodes\n4\n1 +1 0 0\n2 0 1 0\n3 0 0 1\n4 1e-3 1E3 -0\n$EndNodes\n$Elements\n3\n7 8 9 1 2 3 4\n1 2 3 4\n3 1 2 4 3",
}


def make_meshfile_golden():
    read_meshfile = reference_function("simulation.py", "read_meshfile")
    out = {}
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        from jointforces_b200.mesh import spherical_inclusion
        with contextlib.redirect_stdout(io.StringIO()):
            spherical_inclusion(os.path.join(tmp, "gen.msh"), 100, 10000, 0.667)
        texts = dict(MESH_TEXTS)
        texts["generated"] = open(os.path.join(tmp, "gen.msh")).read()
        for name, text in texts.items():
            path = os.path.join(tmp, name + ".msh")
            with open(path, "w") as f:
                f.write(text)
            with contextlib.redirect_stdout(io.StringIO()):
                coords, tets, ri, ro = read_meshfile(path, 7.0, 9.0)
            out[name + "_text"] = np.array(text)
            out[name + "_coords"] = coords
            out[name + "_tets"] = tets
            out[name + "_radii"] = np.array([ri, ro])
    np.savez_compressed(os.path.join(HERE, "reference_meshfile.npz"), **out)
    print("wrote reference_meshfile.npz:", {k: v.shape for k, v in out.items() if not k.endswith("_text")})


def synthetic_series(n_frames=5, grid=40, seed=3):
    """A PIV series like piv.py writes it: per frame a dict x, y, u, v (accumulated deformation on a pixel grid,
    NaN inside the spheroid mask) and a segmentation dict with radius and centroid."""
    rng = np.random.default_rng(seed)
    xs = np.arange(16, 16 + 32 * grid, 32, dtype=np.float64)
    x, y = np.meshgrid(xs, xs)
    series = []
    for f in range(n_frames):
        cx, cy, r = 660.0 + 3.0 * f, 655.0 - 2.0 * f, 88.0 + 1.5 * f
        dx, dy = x - cx, y - cy
        dist = np.sqrt(dx ** 2 + dy ** 2)
        amp = 0.9 * (f + 1) * r * (r / np.maximum(dist, r)) ** 1.2 * (1.0 + 0.4 * np.cos(np.arctan2(dy, dx)) ** 2)
        u = -amp * dx / dist + rng.normal(0.0, 0.15 * (f + 1), x.shape)
        v = -amp * dy / dist + rng.normal(0.0, 0.15 * (f + 1), x.shape)
        u[dist < r] = np.nan
        v[dist < r] = np.nan
        v[3, 5] = np.nan                      # NaN in v only: kept by the u-mask, dropped by the filter
        u[7, 9], v[7, 9] = 0.0, 0.0           # 0/0 in the filter ratio
        series.append(({'x': x, 'y': y, 'u': u, 'v': v}, {'mask': dist < r, 'radius': r, 'centroid': (cx, cy)}))
    return series


def make_lookup_golden():
    from scipy.interpolate import LinearNDInterpolator
    import pandas as pd
    table = np.load("/root/reference/docs/data/Collagen12_BatchA.npy", allow_pickle=True).item()
    create_lookup_functions = reference_function("simulation.py", "create_lookup_functions", {"LinearNDInterpolator": LinearNDInterpolator})
    infer_pressure = reference_function("force.py", "infer_pressure")
    with np.errstate(all="ignore"):
        get_displacement, get_pressure = create_lookup_functions(table)
        rng = np.random.default_rng(11)
        # table queries: inside, on grid points, outside the hull
        qd = np.concatenate([np.exp(rng.uniform(np.log(0.8), np.log(60.0), 4000)), table['x'][::7], [1.0, 50.0, 0.5, 1e3]])
        qp = np.concatenate([np.exp(rng.uniform(np.log(0.05), np.log(2e4), 4000)), np.resize(table['pressure'][::11], len(table['x'][::7])),
                             [0.1, 1e4, 1.0, 1.0]])
        disp = get_displacement(qd, qp)
        qu = np.where(np.isnan(disp), 1e-3, disp) * np.exp(rng.normal(0.0, 0.3, disp.shape))
        qu[::50] *= -1.0
        pres = get_pressure(qd, qu)
        out = {"table_pressure": table['pressure'], "table_x": table['x'], "table_y": table['y'],
               "q_distance": qd, "q_pressure": qp, "displacement": disp, "q_displacement": qu, "pressure": pres}
        # infer_pressure on every frame of a synthetic series, with and without the angle filter
        series = synthetic_series()
        for f, (dis, seg) in enumerate(series):
            for tag, af in (("f20", 20), ("f45", 45.5), ("none", None)):
                res = infer_pressure(np.ravel(dis['x']), np.ravel(dis['y']), np.ravel(dis['u']), np.ravel(dis['v']),
                                     seg['centroid'][0], seg['centroid'][1], series[0][1]['radius'], get_pressure, angle_filter=af)
                for name, arr in zip(("distance", "displacement", "angle", "pressure"), res):
                    out["infer_%d_%s_%s" % (f, tag, name)] = np.asarray(arr, dtype=np.float64)
        # the reference's reconstruct on that series written as files (natsort -> sorted and tqdm -> identity: the file
        # names are zero padded; Excel output disabled, openpyxl is not installed)
        import glob as globmod
        import tempfile
        import yaml  # noqa: F401  (imported inside reconstruct)
        load = reference_function("utils.py", "load")
        load_lookup_functions = reference_function("simulation.py", "load_lookup_functions",
                                                   {"create_lookup_functions": create_lookup_functions})
        reconstruct = reference_function("force.py", "reconstruct", {
            "natsorted": sorted, "glob": globmod.glob, "tqdm": lambda it: it, "load": load, "os": os, "pd": pd,
            "load_lookup_functions": load_lookup_functions, "infer_pressure": infer_pressure})
        pd.DataFrame.to_excel = lambda self, *a, **k: None
        for tag, kw in (("default", {}), ("rmax", {"r_min": 1.5, "r_max": 6.0, "angle_filter": None}),
                        ("continuous", {"continuous_radii": True, "r_min": None})):
            with tempfile.TemporaryDirectory() as tmp:
                np.save(os.path.join(tmp, "table.npy"), table)
                for f, (dis, seg) in enumerate(series):
                    np.save(os.path.join(tmp, "def%06d.npy" % f), dis)
                    np.save(os.path.join(tmp, "seg%06d.npy" % f), seg)
                with contextlib.redirect_stdout(io.StringIO()):
                    df = reconstruct(tmp, os.path.join(tmp, "table.npy"), 1.29, **kw)
                out["reconstruct_%s" % tag] = df.to_numpy(dtype=np.float64)
        out["reconstruct_columns"] = np.array(list(df.columns))
    np.savez_compressed(os.path.join(HERE, "reference_lookup.npz"), **out)
    print("wrote reference_lookup.npz: %d arrays; NaN pressures %d of %d; reconstruct\n%s" % (
        len(out), np.isnan(pres).sum(), pres.size, out["reconstruct_default"][:, :3]))


class RecordingSolver:
    """Stands in for saenopy.Solver: records what jointforces hands over, solves nothing.  `save` stores the nodes
    and a displacement field chosen by the generator so that the reference's curve extraction has data to read."""
    instances = []
    field = None            # callable(nodes, pressure) -> (N,3) displacements written by save()

    def __init__(self):
        self.calls = {}
        RecordingSolver.instances.append(self)

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
        if RecordingSolver.field is not None:
            with open(path, "wb") as f:
                np.savez(f, nodes=self.calls["nodes"], displacements=RecordingSolver.field(self))


class _LoadedSolver:
    def __init__(self, path):
        class _Mesh:
            pass
        z = np.load(path)
        self.mesh = _Mesh()
        self.mesh.nodes, self.mesh.displacements = z["nodes"], z["displacements"]


def make_simulation_golden():
    """spherical_contraction_solver (simulation.py:81-168: boundary conditions, parameters.txt, the calls into the
    solver object), distribute_solver (:171-249: values file, folder names, positional call), extract_deformation_
    curve_solver and create_lookup_table_solver (:263-323), materials.py -- all the reference's own code, with the
    saenopy object replaced by a recorder."""
    import glob as globmod
    import tempfile
    from jointforces_b200.mesh import spherical_inclusion
    out = {}
    mats = {}
    exec(compile(open(os.path.join(REF, "materials.py")).read(), os.path.join(REF, "materials.py"), "exec"), mats)
    for name in ("collagen06", "collagen12", "collagen24", "fibrin40", "matrigel100"):
        out["material_" + name] = np.array([mats[name][k] for k in ("K_0", "D_0", "L_S", "D_S")], dtype=np.float64)
    out["material_linear_394"] = np.array([mats["linear"](394)[k] for k in ("K_0", "D_0", "L_S", "D_S")], dtype=np.float64)

    read_meshfile = reference_function("simulation.py", "read_meshfile")
    g = {"read_meshfile": read_meshfile, "os": os, "Solver": RecordingSolver, "SemiAffineFiberMaterial": lambda *a: tuple(a)}
    solver_fn = reference_function("simulation.py", "spherical_contraction_solver", g)
    with tempfile.TemporaryDirectory() as tmp, contextlib.redirect_stdout(io.StringIO()):
        # (1) one solve on a small mesh, every argument non-default
        spherical_inclusion(os.path.join(tmp, "small.msh"), 100, 10000, 0.667)
        out["small_msh"] = np.array(open(os.path.join(tmp, "small.msh")).read())
        n_small = len(read_meshfile(os.path.join(tmp, "small.msh"))[0])
        U0 = np.random.default_rng(1).normal(0, 1e-7, (n_small, 3))
        RecordingSolver.instances.clear()
        solver_fn(os.path.join(tmp, "small.msh"), os.path.join(tmp, "one"), 123.456, mats["collagen12"], None, None, False, U0, 55, 0.01, 0.02) \
            if False else solver_fn(os.path.join(tmp, "small.msh"), os.path.join(tmp, "one"), 123.456, mats["collagen12"], None, None,
                                    False, U0, 55, 0.01, 0.02)
        rec = RecordingSolver.instances[-1].calls
        out["one_material"] = np.array(rec["material"], dtype=np.float64)
        out["one_nodes"], out["one_tets"] = rec["nodes"], rec["tets"]
        out["one_bc_displacement"], out["one_bc_forces"], out["one_initial"] = rec["bc_displacement"], rec["bc_forces"], rec["initial"]
        out["one_solve_kwargs"] = np.array(repr(sorted((k, v if k != "relrecname" else os.path.basename(v)) for k, v in rec["solve"].items())))
        out["one_save"] = np.array(os.path.basename(rec["save"]))
        out["one_parameters"] = np.array(open(os.path.join(tmp, "one", "parameters.txt"), encoding="utf-8").read())
        out["one_U0"] = U0

        # (2) a sweep through distribute_solver with n_cores=1 on a finer mesh; the recorder's save() writes a smooth
        # radial field scaled with the load so that curve extraction and table building have something to bin
        spherical_inclusion(os.path.join(tmp, "mid.msh"), 100, 10000, 0.1)
        g2 = dict(g)
        g2["spherical_contraction_solver"] = solver_fn
        g2["load_solver"] = _LoadedSolver
        g2["sleep"] = lambda s: None
        dist_fn = reference_function("simulation.py", "distribute_solver", g2)

        def field(solver):
            nodes = solver.calls["nodes"]
            r = np.sqrt(np.sum(nodes ** 2., axis=1))
            load = np.nanmax(np.abs(solver.calls["bc_forces"]))
            amp = 2e-6 * (load / 1e-10) ** 0.7 * (1e-4 / r) ** 1.8 * (1.0 + 0.3 * np.sin(7.0 * nodes[:, 0] / r))
            return -nodes / r[:, None] * amp[:, None]

        RecordingSolver.field = field
        RecordingSolver.instances.clear()
        seen = []
        const_args = {"meshfile": os.path.join(tmp, "mid.msh"), "outfolder": os.path.join(tmp, "sweep"), "material": mats["collagen12"],
                      "max_iter": 77}
        dist_fn("ignored", const_args, var_arg="pressure", start=0.1, end=10000, n=7, log_scaling=True, n_cores=1,
                callback=lambda i, n: seen.append((i, n)))
        out["sweep_const_args_after"] = np.array(repr(sorted(k for k in const_args)))
        out["sweep_values_txt"] = np.array(open(os.path.join(tmp, "sweep", "pressure-values.txt")).read())
        out["sweep_listing"] = np.array(sorted(os.listdir(os.path.join(tmp, "sweep"))))
        out["sweep_callbacks"] = np.array(seen)
        out["sweep_solve_kwargs"] = np.array(repr(sorted((k, v if k != "relrecname" else os.path.basename(v))
                                                         for k, v in RecordingSolver.instances[-1].calls["solve"].items())))
        out["sweep_parameters"] = np.array([open(os.path.join(tmp, "sweep", "simulation%06d" % i, "parameters.txt"), encoding="utf-8").read()
                                            for i in range(7)])
        out["sweep_nodes"] = RecordingSolver.instances[-1].calls["nodes"]
        out["sweep_displacements"] = np.array([np.load(os.path.join(tmp, "sweep", "simulation%06d" % i, "solver.saenopy"))["displacements"]
                                               for i in range(7)])
        out["sweep_msh"] = np.array(open(os.path.join(tmp, "mid.msh")).read())
        # linear branch of the value generator (it spaces the logarithms linearly, :192)
        const_args = {"meshfile": os.path.join(tmp, "small.msh"), "outfolder": os.path.join(tmp, "lin"), "material": mats["collagen12"]}
        RecordingSolver.field = None
        dist_fn("ignored", const_args, start=1, end=100, n=3, log_scaling=False, n_cores=1)
        out["lin_values_txt"] = np.array(open(os.path.join(tmp, "lin", "pressure-values.txt")).read())

        # (3) the reference's curve extraction and table on those folders
        g3 = {"load_solver": _LoadedSolver, "glob": globmod.glob, "tqdm": lambda it: it}
        g3["extract_deformation_curve_solver"] = reference_function("simulation.py", "extract_deformation_curve_solver", g3)
        table_fn = reference_function("simulation.py", "create_lookup_table_solver", g3)
        with np.errstate(all="ignore"), __import__("warnings").catch_warnings():
            __import__("warnings").simplefilter("ignore")
            table = table_fn(os.path.join(tmp, "sweep"), 1, 50, 100)
        order = np.argsort(table["pressure"])
        out["table_pressure"], out["table_x"], out["table_y"] = table["pressure"][order], table["x"], table["y"][order]
    np.savez_compressed(os.path.join(HERE, "reference_simulation.npz"), **out)
    print("wrote reference_simulation.npz: %d arrays, sweep mesh %d nodes, populated bins %d of 100\n%s" % (
        len(out), len(out["sweep_nodes"]), (~np.isnan(out["table_y"][0])).sum(), str(out["one_parameters"])))


if __name__ == "__main__":
    make_meshfile_golden()
    make_lookup_golden()
    make_simulation_golden()
