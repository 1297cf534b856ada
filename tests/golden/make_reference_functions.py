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


if __name__ == "__main__":
    make_meshfile_golden()
    make_lookup_golden()
