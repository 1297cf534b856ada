"""Synthetic spherical-inclusion meshes in the reference's input format.

The reference builds its mesh with the gmsh SDK (``/root/reference/jointforces/mesh.py:5-91``:
two OpenCascade spheres, boolean cut, Delaunay 3D, written as Gmsh 2.2 ASCII with a
``$Jointforces`` trailer at ``mesh.py:86``).  gmsh is not available here and the reference's own
mesh file is a missing blob, so this module generates a *synthetic* member of the same family
(graded element size, h proportional to r) and writes exactly the file layout
``simulation.read_meshfile`` parses (``/root/reference/jointforces/simulation.py:25-78``).

Two generators:

``method='delaunay'`` (default; the family SURVEY.md §8d sized the BASELINE configs with)
    Fibonacci-lattice points on concentric shells, r_k = r_inner * q**k, spacing h = c*r with
    c = (2*pi/6)*lf_iso, shell ratio q ~ 1 + 0.8*c, each shell rotated by 0.7 rad * k about z,
    ``scipy.spatial.Delaunay``, tets with all four nodes on the inner shell dropped.

``method='icoshell'``
    Subdivided icosahedron on every shell, prisms between shells split into three tets by the
    global-index rule (conforming).  No qhull; fast and reproducible to the bit.

Calibration between the reference's gmsh ``length_factor`` and this family:
``lf_iso = (10/3) * length_factor`` -- a convention, not gmsh truth; always
report the actual node and tet counts next to every number.
"""
from __future__ import annotations

import numpy as np

LF_CALIBRATION = 10.0 / 3.0


def _fibonacci_sphere(n: int) -> np.ndarray:
    i = np.arange(n) + 0.5
    z = 1.0 - 2.0 * i / n
    phi = np.pi * (3.0 - np.sqrt(5.0)) * np.arange(n)
    rho = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    return np.stack([rho * np.cos(phi), rho * np.sin(phi), z], axis=1)


def _shell_radii(r_inner: float, r_outer: float, c: float) -> np.ndarray:
    q = 1.0 + 0.8 * c
    k = max(1, int(np.ceil(np.log(r_outer / r_inner) / np.log(q) - 1e-9)))
    return r_inner * (r_outer / r_inner) ** (np.arange(k + 1) / k)


def _delaunay_mesh(r_inner, r_outer, lf_iso):
    from scipy.spatial import Delaunay
    c = (2.0 * np.pi / 6.0) * lf_iso
    radii = _shell_radii(r_inner, r_outer, c)
    n_shell = max(12, int(round(4.0 * np.pi / (0.5 * np.sqrt(3.0) * c * c))))
    unit = _fibonacci_sphere(n_shell)
    pts = []
    for k, r in enumerate(radii):
        a = 0.7 * k
        rot = np.array([[np.cos(a), -np.sin(a), 0.0], [np.sin(a), np.cos(a), 0.0], [0.0, 0.0, 1.0]])
        pts.append(r * unit @ rot.T)
    pts = np.concatenate(pts)
    tri = Delaunay(pts)
    tets = tri.simplices.astype(np.int64)
    inner = np.arange(len(pts)) < n_shell
    tets = tets[~np.all(inner[tets], axis=1)]
    # drop exactly flat cells (qhull may emit them on cospherical input)
    B = pts[tets[:, 1:]] - pts[tets[:, :1]]
    vol = np.abs(np.linalg.det(B)) / 6.0
    h = np.linalg.norm(B, axis=2).max(axis=1)
    tets = tets[vol > 1e-9 * h ** 3]
    return pts, tets


def _icosphere(n_sub: int):
    """Vertices/faces of an icosahedron with every edge split into ``n_sub`` segments."""
    t = (1.0 + np.sqrt(5.0)) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t],
                  [0, -1, -t], [0, 1, -t], [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    v /= np.linalg.norm(v, axis=1)[:, None]
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4],
                  [11, 10, 2], [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8],
                  [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]])
    n = n_sub
    verts, faces = [], []
    key = {}

    def vid(p):
        k = tuple(np.round(p, 9))
        if k not in key:
            key[k] = len(verts)
            verts.append(p)
        return key[k]

    for a, b, c in f:
        idx = {}
        for i in range(n + 1):
            for j in range(n + 1 - i):
                p = (v[a] * (n - i - j) + v[b] * i + v[c] * j) / n
                idx[(i, j)] = vid(p / np.linalg.norm(p))
        for i in range(n):
            for j in range(n - i):
                faces.append([idx[(i, j)], idx[(i + 1, j)], idx[(i, j + 1)]])
                if i + j < n - 1:
                    faces.append([idx[(i + 1, j)], idx[(i + 1, j + 1)], idx[(i, j + 1)]])
    return np.array(verts), np.array(faces, dtype=np.int64)


def _icoshell_mesh(r_inner, r_outer, lf_iso):
    c = (2.0 * np.pi / 6.0) * lf_iso
    # edge length of the subdivided icosahedron on the unit sphere ~ 1.05/n_sub... pick n_sub so h ~ c
    n_sub = max(1, int(round(1.1071 / c)))
    unit, faces = _icosphere(n_sub)
    nv = len(unit)
    radii = _shell_radii(r_inner, r_outer, c)
    pts = np.concatenate([r * unit for r in radii])
    tets = []
    for k in range(len(radii) - 1):
        lo = faces + k * nv
        hi = lo + nv
        # sort each triangle's vertices by global id so shared quads are cut by the same diagonal
        order = np.argsort(lo, axis=1)
        a = np.take_along_axis(lo, order, 1)
        b = np.take_along_axis(hi, order, 1)
        tets.append(np.stack([a[:, 0], a[:, 1], a[:, 2], b[:, 2]], 1))
        tets.append(np.stack([a[:, 0], a[:, 1], b[:, 2], b[:, 1]], 1))
        tets.append(np.stack([a[:, 0], b[:, 1], b[:, 2], b[:, 0]], 1))
    return pts, np.concatenate(tets)


def generate_spherical_inclusion(r_inner=100.0, r_outer=10000.0, length_factor=0.05, lf_iso=None,
                                 method="delaunay"):
    """Return ``(coords_m (N,3) float64 [metres], tets (T,4) int64)``; radii are in micrometres
    like the reference's arguments (``mesh.py:5``), coordinates in metres like its gmsh model
    (``mesh.py:65-66`` multiplies by 1e-6)."""
    if lf_iso is None:
        lf_iso = LF_CALIBRATION * length_factor
    if method == "delaunay":
        pts, tets = _delaunay_mesh(float(r_inner), float(r_outer), float(lf_iso))
    elif method == "icoshell":
        pts, tets = _icoshell_mesh(float(r_inner), float(r_outer), float(lf_iso))
    else:
        raise ValueError("unknown mesh method %r" % (method,))
    # every node must belong to a tet (orphans would give empty stiffness rows)
    used = np.zeros(len(pts), dtype=bool)
    used[tets.ravel()] = True
    if not used.all():
        remap = np.cumsum(used) - 1
        pts = pts[used]
        tets = remap[tets]
    return pts * 1e-6, tets


def write_msh22(outfile, coords_m, tets, r_inner, r_outer, length_factor):
    """Gmsh 2.2 ASCII with only what ``read_meshfile`` reads: ``$Nodes`` (id x y z, ids 1..N in
    file order), ``$Elements`` (tets only: id 4 2 tag tag n1 n2 n3 n4; the parser takes the last
    four tokens of every line, ``simulation.py:72-76``) and the ``$Jointforces`` trailer in the
    exact form of ``mesh.py:86``."""
    n, t = len(coords_m), len(tets)
    with open(outfile, "w") as f:
        f.write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n%d\n" % n)
        ids = np.arange(1, n + 1)
        lines = ["%d %.17g %.17g %.17g" % (i, x, y, z) for i, (x, y, z) in zip(ids, coords_m)]
        f.write("\n".join(lines))
        f.write("\n$EndNodes\n$Elements\n%d\n" % t)
        t1 = np.asarray(tets, dtype=np.int64) + 1
        lines = ["%d 4 2 4 3 %d %d %d %d" % (i + 1, a, b, c, d) for i, (a, b, c, d) in enumerate(t1)]
        f.write("\n".join(lines))
        f.write("\n$EndElements\n")
        f.write('$Jointforces\ninfo=This mesh was created with JOINTFORCES.\ntype=spherical_inclusion\n'
                'r_inner={}\nr_outer={}\nlength_factor={}\n$EndJointforces\n'.format(r_inner, r_outer, length_factor))


def spherical_inclusion(outfile, r_inner=100, r_outer=10000, length_factor=0.05, lf_iso=None,
                        method="delaunay"):
    """Same call as the reference's ``jf.mesh.spherical_inclusion(outfile, r_inner, r_outer,
    length_factor)`` (``mesh.py:5``); writes a synthetic mesh (see module docstring)."""
    coords, tets = generate_spherical_inclusion(r_inner, r_outer, length_factor, lf_iso, method)
    write_msh22(outfile, coords, tets, r_inner, r_outer, length_factor)
    print('Saved spherical inclusion mesh at: {}'.format(outfile))
    return coords, tets
