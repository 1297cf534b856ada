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


if __name__ == "__main__":
    make_meshfile_golden()
