"""The solver-object protocol jointforces drives, backed by the CUDA library.

jointforces never touches finite-element arithmetic itself; it calls nine methods of
``saenopy.Solver`` (``/root/reference/jointforces/simulation.py:100-104,143,146,167-168``) and
reads ``.mesh.nodes / .mesh.displacements / .mesh.forces`` (``:255-258,280-281``; legacy ``.R``,
``.U``, ``:213,286-287``).  This module offers the same object surface -- same method names,
argument meaning, NaN-coded boundary conditions and error behaviour -- and hands the work to
``libjf_b200`` through the C ABI (``include/jf_b200.h``).  There is no CPU path here: without the
built library or without a CUDA device ``solve_boundarycondition`` raises.
"""
from __future__ import annotations

import os
import shutil
import zipfile

import numpy as np

from . import _lib


def build_beams(n=300):
    """Direction sphere of the semi-affine model (saenopy ``build_beams``):
    latitude rings at theta = 2*pi*i/N with round(N sin theta) points each; 286 unit vectors
    for the default n=300."""
    n_ring = int(np.floor(np.sqrt(n * np.pi + 0.5)))
    rows = []
    for i in range(n_ring):
        theta = (2 * np.pi / n_ring) * i
        count = int(np.floor(n_ring * np.sin(theta) + 0.5))
        for j in range(count):
            rho = (2 * np.pi / count) * j
            rows.append((np.sin(theta) * np.cos(rho), np.sin(theta) * np.sin(rho), np.cos(theta)))
    return np.array(rows, dtype=np.float64)


class SemiAffineFiberMaterial:
    """``SemiAffineFiberMaterial(k, d0, lambda_s, ds)`` as constructed at ``simulation.py:101``.

    Holds parameters only; w, w', w'' (three regimes of ``strain.py:336-338``) are evaluated in
    closed form inside the element kernel.  ``None`` switches a regime off like saenopy does."""

    def __init__(self, k, d0=None, lambda_s=None, ds=None):
        self.k = float(k)
        self.d0 = 1e30 if d0 is None else float(d0)
        self.lambda_s = 1e30 if lambda_s is None else float(lambda_s)
        self.ds = 1e30 if ds is None else float(ds)
        self.parameters = dict(k=self.k, d0=self.d0, lambda_s=self.lambda_s, ds=self.ds)

    def as_tuple(self):
        return self.k, self.d0, self.lambda_s, self.ds

    def serialize(self):
        """saenopy's form: ``['SemiAffineFiberMaterial', k, d_0, lambda_s, d_s]``."""
        return ["SemiAffineFiberMaterial", self.k, self.d0, self.lambda_s, self.ds]

    def __repr__(self):
        return "SemiAffineFiberMaterial(k=%r, d0=%r, lambda_s=%r, ds=%r)" % self.as_tuple()


class Mesh:
    """Attribute bag with saenopy's field names."""

    def __init__(self):
        self.nodes = None
        self.tetrahedra = None
        self.displacements = None
        self.displacements_target = None
        self.forces = None
        self.forces_target = None
        self.movable = None
        self.cell_energy = None
        self.strain_energy = 0.0


class Solver:
    """Drop-in for the subset of ``saenopy.Solver`` that jointforces uses."""

    def __init__(self, device=0):
        self.mesh = Mesh()
        self.material_model = None
        self.material_parameters = None
        self.beams = build_beams(300)
        self.device = int(device)
        self.relrec = None
        self.cg_iterations = None
        self.stats = None
        self.info = None

    # ---- setup (simulation.py:102-104) -------------------------------------------------------------
    def set_beams(self, beams=300):
        self.beams = build_beams(beams) if isinstance(beams, (int, np.integer)) else np.asarray(beams, dtype=np.float64)

    def set_material_model(self, material, generate_lookup=True):
        if not hasattr(material, "as_tuple"):
            raise TypeError("material must be a SemiAffineFiberMaterial")
        self.material_model = material
        self.material_parameters = material.serialize()

    def set_nodes(self, data):
        data = np.asarray(data, dtype=np.float64)
        if data.ndim != 2 or data.shape[1] != 3:
            raise ValueError("nodes must have shape (N, 3), got %r" % (data.shape,))
        m = self.mesh
        m.nodes = data.copy()
        n = len(data)
        m.displacements = np.zeros((n, 3))
        m.forces = np.zeros((n, 3))
        m.forces_target = np.zeros((n, 3))
        m.movable = np.ones(n, dtype=bool)

    def set_tetrahedra(self, data):
        data = np.asarray(data)
        if data.ndim != 2 or data.shape[1] != 4:
            raise ValueError("tetrahedra must have shape (T, 4), got %r" % (data.shape,))
        # the reference passes a float array (simulation.py:71-76)
        tets = data.astype(np.int64)
        if self.mesh.nodes is not None and len(tets) and (tets.min() < 0 or tets.max() >= len(self.mesh.nodes)):
            raise ValueError("tetrahedra reference nodes outside [0, N)")
        self.mesh.tetrahedra = tets
        self.mesh.cell_energy = np.zeros(len(tets))

    # ---- boundary conditions (simulation.py:127-146) -------------------------------------
    def set_boundary_condition(self, displacements=None, forces=None):
        m = self.mesh
        if m.nodes is None:
            raise RuntimeError("set_nodes must be called before set_boundary_condition")
        n = len(m.nodes)
        m.displacements = np.zeros((n, 3))
        if displacements is None:
            m.movable = np.ones(n, dtype=bool)
            m.displacements_target = np.full((n, 3), np.nan)
        else:
            displacements = np.asarray(displacements, dtype=np.float64)
            if displacements.shape != (n, 3):
                raise ValueError("displacement boundary condition must have shape (N, 3)")
            m.movable = np.any(np.isnan(displacements), axis=1)
            m.displacements[~m.movable] = displacements[~m.movable]
            m.displacements_target = displacements.copy()
        m.forces_target = np.zeros((n, 3))
        if forces is not None:
            forces = np.asarray(forces, dtype=np.float64)
            if forces.shape != (n, 3):
                raise ValueError("force boundary condition must have shape (N, 3)")
            if np.any(np.isnan(forces[m.movable])):
                raise ValueError("free nodes need a force target (NaN found)")
            m.forces_target[m.movable] = forces[m.movable]

    def set_initial_displacements(self, displacements):
        displacements = np.asarray(displacements, dtype=np.float64)
        m = self.mesh
        if displacements.shape != m.displacements.shape:
            raise ValueError("initial displacements must have shape (N, 3)")
        m.displacements[m.movable] = displacements[m.movable]

    # ---- the hot call (simulation.py:167) -------------------------------------------------------------
    def solve_boundarycondition(self, step_size=0.066, max_iterations=300, rel_conv_crit=0.01, relrecname=None,
                                verbose=False, callback=None):
        m = self.mesh
        if m.nodes is None or m.tetrahedra is None:
            raise RuntimeError("nodes and tetrahedra must be set before solving")
        if self.material_model is None:
            raise RuntimeError("material model must be set before solving")
        if relrecname is None and callback is None and not verbose:
            out = _lib.solve_boundarycondition(m.nodes, m.tetrahedra, m.movable, m.forces_target, m.displacements,
                                               self.material_model.as_tuple(), step_size, max_iterations, rel_conv_crit,
                                               beams=self.beams, device=self.device)
        else:
            out = self._solve_with_progress(step_size, max_iterations, rel_conv_crit, relrecname, verbose, callback)
        m.displacements = out["U"]
        m.forces = out["forces"]
        m.strain_energy = float(out["relrec"][-1, 1])
        self.relrec = out["relrec"]
        self.cg_iterations = out["cg_iters"]
        self.stats = out["stats"]
        return self.relrec

    def _solve_with_progress(self, step_size, max_iterations, rel_conv_crit, relrecname, verbose, callback):
        """saenopy rewrites ``relrecname`` and calls ``callback(self, relrec, i, max_iterations)`` after EVERY Newton
        step (a killed solve leaves its record behind; progress bars move).  The device logs the record rows into
        mapped memory while the host keeps enqueueing steps; this thread follows the rows and does both."""
        m = self.mesh
        done = [0]

        def on_rows(_sys, rows):
            self.relrec = rows
            if relrecname is not None:
                np.savetxt(relrecname, rows)
            for i in range(max(done[0], 1), len(rows)):
                if verbose:
                    print("Newton ", i - 1, ": du=", rows[i, 0], "  Energy=", rows[i, 1], "  Residuum=", rows[i, 2])
                if callback is not None:
                    callback(self, rows[: i + 1], i - 1, max_iterations)
            done[0] = len(rows)
        with _lib.Context(m.nodes, np.ascontiguousarray(m.tetrahedra, dtype=np.int32), m.movable, n_sys=1, beams=self.beams,
                          device=self.device) as ctx:
            ctx.set_material(*self.material_model.as_tuple())
            ctx.set_targets(m.forces_target)
            ctx.set_displacements(m.displacements)
            relrec, cgit, n = ctx.relax(step_size, max_iterations, rel_conv_crit, on_progress=on_rows)
            return dict(U=ctx.displacements(), forces=ctx.forces(), relrec=relrec[0], cg_iters=cgit[0], n_iter=int(n[0]),
                        stats=ctx.stats())

    # ---- persistence (simulation.py:168, 278-291) -------------------------------------------------------
    def save(self, filename, layout="saenopy"):
        save_solver_file(filename, self.mesh.nodes, self.mesh.tetrahedra, self.mesh.displacements, self.mesh.forces,
                         self.mesh.forces_target, self.mesh.movable, self.material_parameters, self.relrec, layout=layout)

    def save_legacy(self, filename):
        """The pre-1.0 archive (keys ``R, T, U, f``) that ``simulation.py:252-258, 289-291`` falls back to."""
        self.save(filename, layout="legacy")

    # legacy attribute names (simulation.py:213, 286-287)
    @property
    def R(self):
        return self.mesh.nodes

    @property
    def T(self):
        return self.mesh.tetrahedra

    @property
    def U(self):
        return self.mesh.displacements

    @property
    def f(self):
        return self.mesh.forces

    def get_contractility(self, center_mode="force"):
        """Sum of the force components pointing at the force centre (used at ``simulation.py:259``)."""
        f = np.nan_to_num(self.mesh.forces)
        R = self.mesh.nodes
        w = np.linalg.norm(f, axis=1)
        if w.sum() == 0:
            return 0.0
        centre = (R * w[:, None]).sum(0) / w.sum() if center_mode == "force" else R.mean(0)
        d = R - centre
        dist = np.linalg.norm(d, axis=1)
        dist[dist == 0] = 1.0
        return float(-np.sum(np.einsum("ij,ij->i", f, d) / dist))


def _material_list(material_parameters):
    """saenopy's ``material.serialize()``: ``['SemiAffineFiberMaterial', k, d_0, lambda_s, d_s]`` (rebuilt on load with
    ``SemiAffineFiberMaterial(*material_parameters[1:])``)."""
    if material_parameters is None or len(material_parameters) == 0:
        return None
    if isinstance(material_parameters, dict):
        m = material_parameters
        return [m.get("type", "SemiAffineFiberMaterial"), m["k"], m["d_0"], m["lambda_s"], m["d_s"]]
    return list(material_parameters)


def save_solver_file(filename, nodes, tets, displacements, forces, forces_target, movable, material_parameters, relrec,
                     layout="saenopy"):
    """Write an npz archive under ``filename`` (for ``solver.saenopy``: no extra suffix).

    ``layout='saenopy'`` (default) follows saenopy's ``Saveable`` as recalled (the package is absent here, DESIGN.md
    "Known gaps"): nested fields flattened to ``group/field`` keys, a list stored as the marker string ``"list"`` under
    its own key plus ``key/0 .. key/n``, and NO key saenopy's ``Solver`` / ``SolverMesh`` constructors do not take --
    ``mesh/{nodes, tetrahedra, displacements, forces, forces_target, movable}``, ``material_parameters`` (+``/0../4``)
    and ``relrec``.  ``layout='legacy'`` writes the pre-1.0 keys ``R, T, U, f`` (plus ``f_target, var``) that
    ``simulation.py:254-258`` reads through its last fallback.  ``load_solver`` below reads both."""
    nodes, displacements = np.asarray(nodes), np.asarray(displacements)
    if layout == "legacy":
        data = {"R": nodes, "T": np.asarray(tets), "U": displacements, "f": np.asarray(forces),
                "f_target": np.asarray(forces_target), "var": np.asarray(movable)}
    elif layout == "saenopy":
        data = {"mesh/nodes": nodes, "mesh/tetrahedra": np.asarray(tets), "mesh/displacements": displacements,
                "mesh/forces": np.asarray(forces), "mesh/forces_target": np.asarray(forces_target),
                "mesh/movable": np.asarray(movable)}
        mat = _material_list(material_parameters)
        if mat is not None:
            data["material_parameters"] = np.asarray("list")
            for i, v in enumerate(mat):
                data["material_parameters/%d" % i] = np.asarray(v)
        if relrec is not None:
            data["relrec"] = np.asarray(relrec)
    else:
        raise ValueError("layout must be 'saenopy' or 'legacy'")
    tmp = str(filename) + ".tmp.npz"
    np.savez(tmp, **data)
    shutil.move(tmp, str(filename))


def load_solver(filename):
    """Counterpart of ``saenopy.solver.load_solver`` for files written by :meth:`Solver.save`
    (``simulation.py:279``), also accepting a legacy ``solver.npz`` with keys R, U, f, T."""
    if not os.path.exists(filename):
        raise FileNotFoundError(filename)
    if not zipfile.is_zipfile(filename):
        raise ValueError("%s is not a solver archive" % filename)
    with np.load(filename, allow_pickle=False) as z:
        keys = set(z.files)
        get = lambda new, old: z[new] if new in keys else (z[old] if old in keys else None)
        s = Solver.__new__(Solver)
        s.mesh = Mesh()
        s.material_model = None
        s.material_parameters = None
        s.beams = None
        s.device = 0
        s.relrec = z["relrec"] if "relrec" in keys else None
        s.cg_iterations = None
        s.stats = None
        s.info = None
        s.mesh.nodes = get("mesh/nodes", "R")
        s.mesh.tetrahedra = get("mesh/tetrahedra", "T")
        s.mesh.displacements = get("mesh/displacements", "U")
        s.mesh.forces = get("mesh/forces", "f")
        s.mesh.forces_target = get("mesh/forces_target", "f_target")
        s.mesh.movable = get("mesh/movable", "var")
        mp = {k.split("/", 1)[1]: z[k][()] for k in keys if k.startswith("material_parameters/")}
        if mp and all(k.isdigit() for k in mp):          # saenopy's list layout: [type, k, d_0, lambda_s, d_s]
            vals = [mp[str(i)] for i in range(len(mp))]
            mp = dict(type=str(vals[0]), k=float(vals[1]), d_0=float(vals[2]), lambda_s=float(vals[3]), d_s=float(vals[4]))
        if mp:
            none_to_big = lambda v: 1e30 if v is None or (isinstance(v, float) and np.isnan(v)) else float(v)
            s.material_model = SemiAffineFiberMaterial(none_to_big(mp["k"]), none_to_big(mp["d_0"]), none_to_big(mp["lambda_s"]),
                                                       none_to_big(mp["d_s"]))
            s.material_parameters = s.material_model.serialize()
    if s.mesh.nodes is None or s.mesh.displacements is None:
        raise ValueError("%s holds neither mesh/nodes nor R" % filename)
    return s
