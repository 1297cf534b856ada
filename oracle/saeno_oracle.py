"""CPU oracle for the saenopy relaxation path that jointforces drives.

TEST INFRASTRUCTURE ONLY.  Nothing in ``jointforces_b200`` may import this module; only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` do, and only as the checker or the timed CPU baseline.

PARITY UNPINNED.  The arithmetic of this path lives in the third-party package ``saenopy``
(``/root/reference/setup.py:15`` pins it only as ``saenopy>=0.7.4``; it is neither vendored under
``/root/reference`` nor installable here).  The reference has no tests and no golden vectors for
this path.  This file therefore restates saenopy's *published algorithm*
 and anchors on the reference's own call sites:

  ``/root/reference/jointforces/simulation.py:100-104``  Solver(), set_material_model, set_nodes, set_tetrahedra
  ``/root/reference/jointforces/simulation.py:127-143``  NaN-coded boundary conditions
  ``/root/reference/jointforces/simulation.py:167``      solve_boundarycondition(step_size, max_iterations, rel_conv_crit, relrecname)
  ``/root/reference/jointforces/strain.py:336-338``      the three-regime stiffness law w''(lambda)
  ``/root/reference/jointforces/simulation.py:707,733``  E = K_0/6 at nu = 0.25 (used by the Lame self-check)

What pins it instead (tests/test_oracle_*.py): finite differences of its own energy, the exact
Lame thick-shell solution in the linear limit, and the early-stop fractions / curve shapes derived
from the reference's shipped lookup tables (tests/golden/, to mesh-discretisation accuracy only).

Everything is float64 numpy + scipy.sparse, structured like saenopy's solver (batches of
tetrahedra, COO->CSR assembly every Newton iteration, plain Hestenes-Stiefel CG with the
squared-norm tolerance, the five-energy stop rule).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as ssp

__all__ = [
    "build_beams", "SemiAffineFiberMaterial", "Mesh", "Solver", "cg",
    "fold_beams", "lame_shell_displacement",
]


# ----------------------------------------------------------------------------------------------
# (A2) direction sphere
# ----------------------------------------------------------------------------------------------
def build_beams(n: int = 300) -> np.ndarray:
    """Unit directions on latitude rings; saenopy ``build_beams``.

    n=300 -> 30 nominal rings, 14 non-empty, 286 directions.
    """
    n = int(np.floor(np.sqrt(n * np.pi + 0.5)))
    beams = []
    for i in range(n):
        theta = (2 * np.pi / n) * i
        jmax = int(np.floor(n * np.sin(theta) + 0.5))
        for j in range(jmax):
            rho = (2 * np.pi / jmax) * j
            beams.append([np.sin(theta) * np.cos(rho), np.sin(theta) * np.sin(rho), np.cos(theta)])
    return np.array(beams, dtype=np.float64)


def fold_beams(beams: np.ndarray, tol: float = 1e-12):
    """Merge antipodal pairs (every term of A5-A8 is even in s).  Returns (dirs, weights).

    For the default table 228 of 286 directions pair up; the two 29-point rings do not, so the
    folded table has 114 + 58 = 172 weighted directions.  Used only to test that folding is an
    identity (the product kernels fold; the oracle itself never does).
    """
    from scipy.spatial import cKDTree
    tree = cKDTree(beams)
    d, j = tree.query(-beams)
    keep, w = [], []
    for i in range(len(beams)):
        if d[i] < tol and j[i] != i:
            if i < j[i]:
                keep.append(i)
                w.append(2.0)
        else:
            keep.append(i)
            w.append(1.0)
    return beams[keep], np.array(w)


# ----------------------------------------------------------------------------------------------
# (A1) material
# ----------------------------------------------------------------------------------------------
def _expm1_minus_x(x):
    """expm1(x) - x without cancellation for small |x| (needed when d0 or ds is 1e30)."""
    x = np.asarray(x, dtype=np.float64)
    small = np.abs(x) < 0.125
    xs = np.where(small, x, 0.0)
    # x^2/2! + ... + x^13/13!  (Horner); truncation < 1e-19 relative for |x| < 1/8
    acc = np.full_like(xs, 1.0 / 6227020800.0)
    for inv in (1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0,
                1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5):
        acc = acc * xs + inv
    series = acc * xs * xs
    xl = np.where(small, 1.0, x)
    direct = np.expm1(xl) - xl
    return np.where(small, series, direct)


class SemiAffineFiberMaterial:
    """w''(lambda) in three regimes exactly as ``strain.py:336-338``; w', w are its integrals
    with w(0)=w'(0)=0.  Positional signature (k, d0, lambda_s, ds) as
    constructed at ``simulation.py:101``.

    ``mode='closed'`` evaluates the integrals in closed form (expm1-safe).  ``mode='lut'``
    emulates saenopy's sampled table (lambda = -1 ... 4, step 1e-6, w'' clipped at 1e11, w' and w
    by cumulative sums re-zeroed at lambda=0).  The lookup rule is recalled, not verified (saenopy is absent):
    ``lut_index='interp'`` (default) is linear interpolation between the samples floor(q) and floor(q)+1 with the
    index clamped two short of the end -- saenopy's ``sample_and_integrate_function.lookup`` as recalled
    [RECALL]; 'ceil' / 'round' / 'floor' read a single sample and are kept to bound the effect of the rule.
    """

    def __init__(self, k, d0=None, lambda_s=None, ds=None, mode="closed", lut_index="interp"):
        self.k = float(k)
        self.d0 = float(d0) if d0 is not None else 1e30
        self.lambda_s = float(lambda_s) if lambda_s is not None else 1e30
        self.ds = float(ds) if ds is not None else 1e30
        self.mode = mode
        self.lut_index = lut_index
        self._lut = None
        self.parameters = dict(k=self.k, d_0=self.d0, lambda_s=self.lambda_s, d_s=self.ds)

    # closed forms ------------------------------------------------------------------------
    def stiffness(self, lam):
        """w''(lambda): k*exp(lam/d0) | k | k*exp((lam-ls)/ds)   (strain.py:336-338)."""
        lam = np.asarray(lam, dtype=np.float64)
        k, d0, ls, ds = self.k, self.d0, self.lambda_s, self.ds
        out = np.full_like(lam, k)
        neg = lam < 0
        out[neg] = k * np.exp(lam[neg] / d0)
        stf = lam >= ls
        out[stf] = k * np.exp((lam[stf] - ls) / ds)
        return out

    def _closed(self, lam):
        lam = np.asarray(lam, dtype=np.float64)
        k, d0, ls, ds = self.k, self.d0, self.lambda_s, self.ds
        # linear regime
        w2 = np.full_like(lam, k)
        w1 = k * lam
        w0 = 0.5 * k * lam * lam
        # buckling regime lam < 0
        neg = lam < 0
        if np.any(neg):
            x = lam[neg] / d0
            em = np.expm1(x)
            w2[neg] = k * (1.0 + em)
            w1[neg] = k * d0 * em
            w0[neg] = k * d0 * d0 * _expm1_minus_x(x)
        # stiffening regime lam >= ls
        stf = lam >= ls
        if np.any(stf):
            y = (lam[stf] - ls) / ds
            em = np.expm1(y)
            w2[stf] = k * (1.0 + em)
            w1[stf] = k * ls + k * ds * em
            w0[stf] = 0.5 * k * ls * ls + k * ls * (lam[stf] - ls) + k * ds * ds * _expm1_minus_x(y)
        return w0, w1, w2

    # sampled table (saenopy emulation) ---------------------------------------------------
    _LMIN, _LMAX, _LSTEP = -1.0, 4.0, 1e-6

    def _build_lut(self):
        lam = np.arange(self._LMIN, self._LMAX, self._LSTEP)
        y = self.stiffness(lam)
        y[y > 10e10] = 10e10
        i0 = int(np.ceil((0.0 - self._LMIN) / self._LSTEP))
        w1 = np.cumsum(y * self._LSTEP)
        w1 -= w1[i0]
        w0 = np.cumsum(w1 * self._LSTEP)
        w0 -= w0[i0]
        self._lut = (w0, w1, y)

    def _lookup(self, lam):
        if self._lut is None:
            self._build_lut()
        q = (np.asarray(lam, dtype=np.float64) - self._LMIN) / self._LSTEP
        if self.lut_index == "interp":
            li = np.floor(q)
            dli = q - li
            last = (self._LMAX - self._LMIN) / self._LSTEP - 2
            at_end = li > last
            li = np.where(at_end, float(int(last)), li)
            dli = np.where(at_end, 0.0, dli)
            lii = np.clip(li.astype(np.int64), 0, len(self._lut[0]) - 2)
            return tuple((1 - dli) * tab[lii] + dli * tab[lii + 1] for tab in self._lut)
        if self.lut_index == "ceil":
            idx = np.ceil(q)
        elif self.lut_index == "round":
            idx = np.round(q)
        else:
            idx = np.floor(q)
        idx = np.clip(idx.astype(np.int64), 0, len(self._lut[0]) - 1)
        return self._lut[0][idx], self._lut[1][idx], self._lut[2][idx]

    def __call__(self, lam):
        """lambda -> (w, w', w'')."""
        if self.mode == "lut":
            return self._lookup(lam)
        return self._closed(lam)


# ----------------------------------------------------------------------------------------------
# mesh container (attribute names as the reference reads them: simulation.py:255-258, 280-281)
# ----------------------------------------------------------------------------------------------
class Mesh:
    def __init__(self):
        self.nodes = None
        self.tetrahedra = None
        self.displacements = None
        self.forces = None
        self.forces_target = None
        self.movable = None
        self.tetrahedra_volumes = None
        self.cell_energy = None
        self.strain_energy = 0.0


# ----------------------------------------------------------------------------------------------
# (A13) conjugate gradient
# ----------------------------------------------------------------------------------------------
def cg(A, b, maxiter=1000, tol=0.00001, return_iterations=False):
    """Plain Hestenes-Stiefel CG from x=0, stop when |r|^2 < tol*|b|^2."""
    normb = float(np.inner(b, b))
    if normb == 0:
        x = np.zeros_like(b)
        return (x, 0) if return_iterations else x
    x = np.zeros_like(b)
    r = b - A @ x
    p = r
    resid = float(np.inner(p, p))
    it = 0
    for it in range(1, maxiter + 1):
        Ap = A @ p
        alpha = resid / np.sum(p * Ap)
        x = x + alpha * p
        r = r - alpha * Ap
        rsnew = float(np.inner(r, r))
        if rsnew < tol * normb:
            break
        beta = rsnew / resid
        p = r + beta * p
        resid = rsnew
    return (x, it) if return_iterations else x


# ----------------------------------------------------------------------------------------------
# solver
# ----------------------------------------------------------------------------------------------
class Solver:
    """Restatement of the saenopy ``Solver`` protocol jointforces calls."""

    BATCH = 1000  # tetrahedra per batch, as saenopy

    def __init__(self):
        self.mesh = Mesh()
        self.material_model = None
        self.s = None
        self.N_b = 0
        self.Phi = None
        self.K_glo = None
        self.relrec = None
        self.cg_iterations = []      # CG iterations taken in each Newton step (diagnostic)
        self.energy_rule = "free_tets"   # E_glo counts tets with >= 1 free node (saenopy [RECALL]); "all" = every tet
        self.set_beams(300)

    # --- setup --------------------------------------------------------------------------------
    def set_beams(self, beams=300):
        if isinstance(beams, (int, np.integer)):
            beams = build_beams(int(beams))
        self.s = np.asarray(beams, dtype=np.float64)
        self.N_b = self.s.shape[0]

    def set_material_model(self, material):
        self.material_model = material

    def set_nodes(self, data):
        data = np.asarray(data, dtype=np.float64)
        assert data.ndim == 2 and data.shape[1] == 3, "nodes must be (N, 3)"
        m = self.mesh
        m.nodes = data.copy()
        n = len(data)
        m.displacements = np.zeros((n, 3))
        m.forces = np.zeros((n, 3))
        m.forces_target = np.zeros((n, 3))
        m.movable = np.ones(n, dtype=bool)

    def set_tetrahedra(self, data):
        # the reference hands over a float array (simulation.py:71-76); cast like saenopy does
        data = np.asarray(data)
        assert data.ndim == 2 and data.shape[1] == 4, "tetrahedra must be (T, 4)"
        self.mesh.tetrahedra = data.astype(np.int64)
        self.mesh.cell_energy = np.zeros(len(data))

    def set_boundary_condition(self, displacements=None, forces=None):
        """NaN-coded: a node is free iff any displacement component is NaN (Appendix A11)."""
        m = self.mesh
        n = len(m.nodes)
        if displacements is None:
            m.movable = np.ones(n, dtype=bool)
            m.displacements = np.zeros((n, 3))
        else:
            displacements = np.asarray(displacements, dtype=np.float64)
            assert displacements.shape == (n, 3)
            m.movable = np.any(np.isnan(displacements), axis=1)
            m.displacements = np.zeros((n, 3))
            m.displacements[~m.movable] = displacements[~m.movable]
        m.forces_target = np.zeros((n, 3))
        if forces is not None:
            forces = np.asarray(forces, dtype=np.float64)
            assert forces.shape == (n, 3)
            m.forces_target[m.movable] = forces[m.movable]

    def set_initial_displacements(self, displacements):
        displacements = np.asarray(displacements, dtype=np.float64)
        m = self.mesh
        m.displacements[m.movable] = displacements[m.movable]

    # --- one-off geometry (A3, A7) --------------------------------------------------------------
    def _compute_phi(self):
        m = self.mesh
        chi = np.array([[-1.0, -1.0, -1.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]])
        T = m.tetrahedra
        B = m.nodes[T[:, 1:4]] - m.nodes[T[:, 0]][:, None, :]
        B = B.transpose(0, 2, 1)                      # columns are the edge vectors
        m.tetrahedra_volumes = np.abs(np.linalg.det(B)) / 6.0
        self.Phi = chi @ np.linalg.inv(B)             # (T, 4, 3)

    def _compute_connections(self):
        m = self.mesh
        T = m.tetrahedra
        nt = len(T)
        # force entries: (node, component) for every (t, m, i)
        self._f_rows = np.repeat(T.ravel(), 3)
        self._f_cols = np.tile(np.arange(3), nt * 4)
        # stiffness entries (t, m, r, i, l) -> (3*T[t,m]+i, 3*T[t,r]+l), rows of free nodes only
        free_row = m.movable[T]                                    # (T, 4)
        filt = np.broadcast_to(free_row[:, :, None, None, None], (nt, 4, 4, 3, 3))
        rows = np.broadcast_to((3 * T)[:, :, None, None, None] + np.arange(3)[None, None, None, :, None], (nt, 4, 4, 3, 3))
        cols = np.broadcast_to((3 * T)[:, None, :, None, None] + np.arange(3)[None, None, None, None, :], (nt, 4, 4, 3, 3))
        self._count_energy = np.any(free_row, axis=1)              # (T,) tets whose energy counts in E_glo
        self._K_filter = filt.ravel().copy()
        self._K_rows = rows.ravel()[self._K_filter]
        self._K_cols = cols.ravel()[self._K_filter]

    def _prepare(self):
        self._compute_phi()
        self._compute_connections()
        self.s_star = self.Phi @ self.s.T                          # (T, 4, N_b)

    # --- (A4)-(A10) element evaluation + assembly -------------------------------------------------
    def element_quantities(self, t=slice(None)):
        """Per-tet energy (T,), force (T,4,3) and stiffness (T,4,4,3,3) for the slice ``t``."""
        m = self.mesh
        s = self.s
        u_t = m.displacements[m.tetrahedra[t]]                     # (t, 4, 3)
        F = np.eye(3) + np.einsum("tmi,tmj->tij", u_t, self.Phi[t])
        s_bar = F @ s.T                                            # (t, 3, b)
        ell = np.linalg.norm(s_bar, axis=1)                        # (t, b)
        w0, w1, w2 = self.material_model(ell - 1.0)
        v_nb = (m.tetrahedra_volumes[t] / self.N_b)[:, None]
        dEdsbar = -(w1 / ell) * v_nb                               # g_b   (A6)
        dEdsbarbar = (ell * w2 - w1) / ell ** 3 * v_nb             # h_b   (A6)
        energy = np.mean(w0, axis=1) * m.tetrahedra_volumes[t]     # (A5)
        sst = self.s_star[t]                                       # (t, 4, b)
        # f_tmi = s*_tmb s'_tib g_tb                                 (A7)
        f_el = np.matmul(sst, (s_bar * dEdsbar[:, None, :]).transpose(0, 2, 1))
        # K_tmril = 1/2 s*_tmb s*_trb (h_tb s'_tib s'_tlb - g_tb d_il)   (A8)
        nt, nb = ell.shape
        q = 0.5 * (dEdsbarbar[:, None, None, :] * s_bar[:, :, None, :] * s_bar[:, None, :, :]
                   - np.eye(3)[None, :, :, None] * dEdsbar[:, None, None, :])          # (t, i, l, b)
        ss = (sst[:, :, None, :] * sst[:, None, :, :]).reshape(nt, 16, nb)              # (t, mr, b)
        K_el = np.matmul(ss, q.reshape(nt, 9, nb).transpose(0, 2, 1)).reshape(nt, 4, 4, 3, 3)
        return energy, f_el, K_el

    def _update_glo_f_and_K(self):
        m = self.mesh
        nt = len(m.tetrahedra)
        f_el = np.zeros((nt, 4, 3))
        K_el = np.zeros((nt, 4, 4, 3, 3))
        m.strain_energy = 0.0
        for i in range(int(np.ceil(nt / self.BATCH))):
            t = slice(i * self.BATCH, (i + 1) * self.BATCH)
            e, f_el[t], K_el[t] = self.element_quantities(t)
            m.cell_energy[t] = e
            # only tetrahedra with at least one free node count (saenopy `_update_energy` / SAENO `countEnergy`
            # [RECALL]); all-fixed tets cannot change.  Zero anyway under simulation.py:127-129 (outer shell at 0).
            m.strain_energy += np.sum(e[self._count_energy[t]]) if self.energy_rule == "free_tets" else np.sum(e)
        n = len(m.nodes)
        m.forces = ssp.coo_matrix((f_el.ravel(), (self._f_rows, self._f_cols)), shape=(n, 3)).toarray()
        self.K_glo = ssp.coo_matrix((K_el.ravel()[self._K_filter], (self._K_rows, self._K_cols)),
                                    shape=(3 * n, 3 * n)).tocsr()
        self._last_f_el, self._last_K_el = f_el, K_el

    # --- (A12) one relaxation step ------------------------------------------------------------------
    def _solve_CG(self, step_size):
        m = self.mesh
        ff = m.forces - m.forces_target
        ff[~m.movable, :] = 0
        uu, it = cg(self.K_glo, ff.ravel(), maxiter=3 * len(m.nodes), tol=0.00001, return_iterations=True)
        self.cg_iterations.append(it)
        uu = uu.reshape(ff.shape)
        m.displacements[m.movable] += uu[m.movable] * step_size
        return float(np.sum(uu[m.movable] ** 2) * step_size * step_size)

    # --- (A14) driver ------------------------------------------------------------------------------------
    def solve_boundarycondition(self, step_size=0.066, max_iterations=300, rel_conv_crit=0.01,
                                relrecname=None, verbose=False, callback=None):
        m = self.mesh
        self._prepare()
        self._update_glo_f_and_K()
        relrec = [[0.0, m.strain_energy, float(np.sum(m.forces[m.movable] ** 2))]]
        self.cg_iterations = []
        for i in range(max_iterations):
            du = self._solve_CG(step_size)
            self._update_glo_f_and_K()
            ff = float(np.sum((m.forces[m.movable] - m.forces_target[m.movable]) ** 2))
            if verbose:
                print("Newton ", i, ": du=", du, "  Energy=", m.strain_energy, "  Residuum=", ff,
                      " cg=", self.cg_iterations[-1])
            relrec.append([du, m.strain_energy, ff])
            if relrecname is not None:
                np.savetxt(relrecname, relrec)
            if callback is not None:
                callback(self, relrec, i, max_iterations)
            if i > 6:
                last_es = np.array(relrec)[:-6:-1, 1]
                e_mean = np.mean(last_es)
                e_std = np.std(last_es) / np.sqrt(5)
                if e_std / e_mean < rel_conv_crit:
                    break
        self.relrec = np.array(relrec)
        return self.relrec

    # --- diagnostics ---------------------------------------------------------------------------------------
    def total_energy(self):
        self._prepare()
        e, _, _ = self.element_quantities()
        return float(np.sum(e))


# ----------------------------------------------------------------------------------------------
# analytic anchor: pressurised cavity in a thick shell fixed at the outer radius
# ----------------------------------------------------------------------------------------------
def lame_shell_displacement(r, pressure, k0, r_inner, r_outer, nu=0.25):
    """|u(r)| of the exact linear-elastic solution: E = K_0/6
    (``simulation.py:707``), G = E/(2(1+nu)) = K_0/15, u(r_outer) = 0, radial traction ``pressure``
    on the cavity wall."""
    E = k0 / 6.0
    G = E / (2 * (1 + nu))
    lam = E * nu / ((1 + nu) * (1 - 2 * nu))
    a, b = r_inner, r_outer
    # u = A r + B / r^2,  sigma_rr = (3 lam + 2 G) A - 4 G B / r^3
    # u(b) = 0 -> A = -B / b^3 ; |sigma_rr(a)| = p
    B = pressure / ((3 * lam + 2 * G) / b ** 3 + 4 * G / a ** 3)
    return np.abs(B * (1.0 / r ** 2 - r / b ** 3))
