"""ctypes binding of libjf_b200.so (include/jf_b200.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is visible when a
compute entry point is called, this raises.  The library is built in-tree by
``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C jointforces_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# JF_B200_LIB points at an instrumented build of the same library (profiling only)
LIB_PATH = os.environ.get("JF_B200_LIB") or os.path.join(_HERE, "_lib", "libjf_b200.so")

JF_FLAG_NO_REORDER = 1
JF_FLAG_NO_FOLD = 2
JF_FLAG_NO_RESIDENT = 4

# every symbol include/jf_b200.h declares (tests check the library exports each of them)
SYMBOLS = [
    "jf_last_error", "jf_version", "jf_device_count", "jf_ctx_create", "jf_ctx_destroy", "jf_set_material",
    "jf_set_targets", "jf_set_displacements", "jf_get_displacements", "jf_restore_displacements", "jf_get_forces", "jf_evaluate", "jf_get_energy",
    "jf_get_element_quantities", "jf_get_stiffness_blocks", "jf_cg", "jf_get_cg_solution", "jf_relax", "jf_get_stats",
    "jf_reset_stats", "jf_get_info", "jf_relax_progress", "jf_solve_boundarycondition", "jf_bench_fp64", "jf_set_curve_bins", "jf_get_curves", "jf_apply_stiffness",
    "jf_msh_open", "jf_msh_read", "jf_msh_close",
    "jf_table_create", "jf_table_destroy", "jf_table_eval", "jf_infer_pressure", "jf_reconstruct_frames", "jf_table_launches",
]


class JFError(RuntimeError):
    pass


class Stats(C.Structure):
    _fields_ = [("ms_element", C.c_double), ("ms_assembly", C.c_double), ("ms_cg", C.c_double), ("ms_setup", C.c_double),
                ("tet_updates", C.c_int64), ("cg_iterations", C.c_int64), ("cg_launches", C.c_int64),
                ("kernel_launches", C.c_int64), ("newton_steps", C.c_int64), ("ms_relax", C.c_double),
                ("bytes_h2d", C.c_int64), ("bytes_d2h", C.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Info(C.Structure):
    _fields_ = [("n_nodes", C.c_int64), ("n_free", C.c_int64), ("n_tets", C.c_int64), ("n_blocks", C.c_int64),
                ("n_slots", C.c_int64), ("n_sys", C.c_int32), ("n_dirs", C.c_int32), ("n_beams", C.c_int32),
                ("sm_count", C.c_int32), ("cg_grid", C.c_int32), ("cg_block", C.c_int32), ("elem_grid", C.c_int32),
                ("elem_block", C.c_int32), ("bytes_device", C.c_int64), ("cg_resident", C.c_int32),
                ("cg_wmax", C.c_int32), ("cg_stream_halo", C.c_int32), ("cg_fused", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_bp = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_vp = C.c_void_p

_lib = None


def load():
    """Load the shared library (once).  Raises JFError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise JFError("libjf_b200.so not found at %s -- build it first (python -c 'import __graft_entry__ as g; "
                      "g.build()'); there is no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.jf_last_error.restype = C.c_char_p
    L.jf_last_error.argtypes = []
    L.jf_version.restype = C.c_int
    L.jf_device_count.restype = C.c_int
    L.jf_ctx_create.argtypes = [C.POINTER(_vp), C.c_int, C.c_int, C.c_int64, _dp, C.c_int64, _ip, _bp, C.c_int, _vp, C.c_uint32]
    L.jf_ctx_destroy.argtypes = [_vp]
    L.jf_ctx_destroy.restype = None
    L.jf_set_material.argtypes = [_vp, C.c_int] + [C.c_double] * 4
    L.jf_set_targets.argtypes = [_vp, C.c_int, _dp]
    L.jf_set_displacements.argtypes = [_vp, C.c_int, _dp]
    L.jf_get_displacements.argtypes = [_vp, C.c_int, _dp]
    L.jf_restore_displacements.argtypes = [_vp, C.c_int]
    L.jf_get_forces.argtypes = [_vp, C.c_int, _dp]
    L.jf_evaluate.argtypes = [_vp]
    L.jf_get_energy.argtypes = [_vp, C.c_int] + [C.POINTER(C.c_double)] * 3
    L.jf_get_element_quantities.argtypes = [_vp, C.c_int, _vp, _vp, _vp]
    L.jf_get_stiffness_blocks.argtypes = [_vp, C.c_int, C.POINTER(C.c_int64), _vp, _vp, _vp]
    L.jf_cg.argtypes = [_vp, C.c_double, C.c_int64, _ip]
    L.jf_get_cg_solution.argtypes = [_vp, C.c_int, _dp]
    L.jf_relax.argtypes = [_vp, C.c_double, C.c_int, C.c_double, C.c_double, _dp, _vp, _ip]
    L.jf_relax_progress.argtypes = [_vp, C.c_int, C.POINTER(C.c_int32), _vp, C.c_int]
    L.jf_get_stats.argtypes = [_vp, C.POINTER(Stats)]
    L.jf_reset_stats.argtypes = [_vp]
    L.jf_get_info.argtypes = [_vp, C.POINTER(Info)]
    L.jf_solve_boundarycondition.argtypes = [C.c_int, C.c_int64, _dp, C.c_int64, _ip, _bp, _dp, _dp, C.c_int, _vp] + \
        [C.c_double] * 4 + [C.c_double, C.c_int, C.c_double, _dp, _vp, _ip, _vp, C.POINTER(Stats)]
    L.jf_bench_fp64.argtypes = [C.c_int, C.POINTER(C.c_double)]
    L.jf_apply_stiffness.argtypes = [_vp, C.c_int, _dp, _dp]
    L.jf_set_curve_bins.argtypes = [_vp, C.c_int, _ip, _vp, C.c_double]
    L.jf_get_curves.argtypes = [_vp, _dp]
    L.jf_msh_open.argtypes = [C.c_char_p, C.POINTER(_vp), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int32),
                              C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.jf_msh_read.argtypes = [_vp, _vp, _vp, _vp, C.c_int32]
    L.jf_msh_close.argtypes = [_vp]
    L.jf_msh_close.restype = None
    _lp = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
    L.jf_table_create.argtypes = [C.POINTER(_vp), C.c_int, C.c_int64, _dp, _dp, C.c_int64, _ip, _dp, _ip]
    L.jf_table_destroy.argtypes = [_vp]
    L.jf_table_destroy.restype = None
    L.jf_table_eval.argtypes = [_vp, C.c_int64, _dp, _dp, C.c_int, C.c_int, C.c_int, _dp]
    L.jf_infer_pressure.argtypes = [_vp, C.c_int64, _dp, _dp, _dp, _dp, C.c_double, C.c_double, C.c_double, C.c_int, C.c_double,
                                    _dp, _dp, _dp, _dp, C.POINTER(C.c_int64)]
    L.jf_reconstruct_frames.argtypes = [_vp, C.c_int32, _lp, _dp, _dp, _dp, _dp, _dp, C.c_int, C.c_double, C.c_double, C.c_double,
                                        _dp, _dp]
    L.jf_table_launches.argtypes = [_vp]
    L.jf_table_launches.restype = C.c_int64
    for name in SYMBOLS:
        fn = getattr(L, name)
        if name not in ("jf_last_error", "jf_ctx_destroy", "jf_msh_close", "jf_table_destroy", "jf_table_launches"):
            fn.restype = C.c_int
    _lib = L
    return L


def check(rc):
    if rc < 0:
        raise JFError("jf_b200 error %d: %s" % (rc, load().jf_last_error().decode("utf-8", "replace")))
    return rc


def device_count():
    return check(load().jf_device_count())


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_vp)


JF_ERR_FORMAT = -6


def read_msh(path, want_int32=False, n_threads=0):
    """Threaded Gmsh 2.2 ASCII reader (jf_msh_*).  Returns ``(coords (N,3) in file units, tets (T,4) 0-based as
    float64 -- or int32 with ``want_int32`` --, radii)`` with ``radii = (r_inner, r_outer)`` in micrometres or None
    when the file has no ``$Jointforces`` block; returns None for files the fast path leaves to the line loop."""
    L = load()
    h = _vp()
    nn, ne, has = C.c_int64(), C.c_int64(), C.c_int32()
    ri, ro = C.c_double(), C.c_double()
    rc = L.jf_msh_open(os.fsencode(path), C.byref(h), C.byref(nn), C.byref(ne), C.byref(has), C.byref(ri), C.byref(ro))
    if rc == JF_ERR_FORMAT:
        return None
    check(rc)
    try:
        coords = np.empty((nn.value, 3), dtype=np.float64)
        tets = np.empty((ne.value, 4), dtype=np.int32 if want_int32 else np.float64)
        rc = L.jf_msh_read(h, _ptr(coords), None if want_int32 else _ptr(tets), _ptr(tets) if want_int32 else None, n_threads)
        if rc == JF_ERR_FORMAT:
            return None
        check(rc)
    finally:
        L.jf_msh_close(h)
    return coords, tets, ((ri.value, ro.value) if has.value else None)


class Context:
    """Thin RAII wrapper over jf_ctx: one mesh + boundary mask, ``n_sys`` load cases."""

    def __init__(self, nodes, tets, movable, n_sys=1, beams=None, device=0, flags=0):
        L = load()
        self._L = L
        self._h = _vp()
        nodes = np.ascontiguousarray(nodes, dtype=np.float64)
        tets = np.ascontiguousarray(tets, dtype=np.int32)
        movable = np.ascontiguousarray(movable, dtype=np.uint8)
        if nodes.ndim != 2 or nodes.shape[1] != 3 or tets.ndim != 2 or tets.shape[1] != 4 or movable.shape != (len(nodes),):
            raise ValueError("nodes (N,3), tets (T,4), movable (N,) expected")
        if beams is not None:
            beams = np.ascontiguousarray(beams, dtype=np.float64)
        self.N, self.T, self.n_sys = len(nodes), len(tets), int(n_sys)
        check(L.jf_ctx_create(C.byref(self._h), int(device), int(n_sys), self.N, nodes, self.T, tets, movable,
                              0 if beams is None else len(beams), _ptr(beams), int(flags)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.jf_ctx_destroy(self._h)
            self._h = _vp()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # --- setters ---------------------------------------------------------------------------------
    def set_material(self, k, d0, lambda_s, ds, sys=-1):
        check(self._L.jf_set_material(self._h, int(sys), float(k), float(d0), float(lambda_s), float(ds)))

    def set_targets(self, f_target, sys=0):
        check(self._L.jf_set_targets(self._h, int(sys), np.ascontiguousarray(f_target, dtype=np.float64)))

    def set_displacements(self, U, sys=0):
        check(self._L.jf_set_displacements(self._h, int(sys), np.ascontiguousarray(U, dtype=np.float64)))

    def restore_displacements(self, sys=-1):
        check(self._L.jf_restore_displacements(self._h, int(sys)))

    # --- getters ---------------------------------------------------------------------------------
    def displacements(self, sys=0):
        out = np.empty((self.N, 3))
        check(self._L.jf_get_displacements(self._h, int(sys), out))
        return out

    def forces(self, sys=0):
        out = np.empty((self.N, 3))
        check(self._L.jf_get_forces(self._h, int(sys), out))
        return out

    def energy(self, sys=0):
        e, r2, f2 = C.c_double(), C.c_double(), C.c_double()
        check(self._L.jf_get_energy(self._h, int(sys), C.byref(e), C.byref(r2), C.byref(f2)))
        return e.value, r2.value, f2.value

    def element_quantities(self, sys=0, want_K=True):
        E, f = np.empty(self.T), np.empty((self.T, 4, 3))
        K = np.empty((self.T, 4, 4, 3, 3)) if want_K else None
        check(self._L.jf_get_element_quantities(self._h, int(sys), _ptr(E), _ptr(f), _ptr(K)))
        return E, f, K

    def stiffness_blocks(self, sys=0):
        n = C.c_int64()
        check(self._L.jf_get_stiffness_blocks(self._h, int(sys), C.byref(n), None, None, None))
        rows, cols = np.empty(n.value, np.int32), np.empty(n.value, np.int32)
        blocks = np.empty((n.value, 3, 3))
        check(self._L.jf_get_stiffness_blocks(self._h, int(sys), C.byref(n), _ptr(rows), _ptr(cols), _ptr(blocks)))
        return rows, cols, blocks

    def stiffness_csr(self, sys=0):
        """Assembled stiffness as scipy CSR (3N x 3N) in the caller's numbering (parity hook)."""
        import scipy.sparse as ssp
        rows, cols, blocks = self.stiffness_blocks(sys)
        r = (3 * rows[:, None, None] + np.arange(3)[None, :, None] + np.zeros((1, 1, 3), int)).ravel()
        c = (3 * cols[:, None, None] + np.arange(3)[None, None, :] + np.zeros((1, 3, 1), int)).ravel()
        return ssp.coo_matrix((blocks.ravel(), (r, c)), shape=(3 * self.N, 3 * self.N)).tocsr()

    # --- compute ---------------------------------------------------------------------------------
    def evaluate(self):
        check(self._L.jf_evaluate(self._h))

    def cg(self, tol=1e-5, maxiter=0):
        it = np.zeros(self.n_sys, np.int32)
        check(self._L.jf_cg(self._h, float(tol), int(maxiter), it))
        return it

    def apply_stiffness(self, x, sys=0):
        out = np.empty((self.N, 3))
        check(self._L.jf_apply_stiffness(self._h, int(sys), np.ascontiguousarray(x, dtype=np.float64), out))
        return out

    def cg_solution(self, sys=0):
        out = np.empty((self.N, 3))
        check(self._L.jf_get_cg_solution(self._h, int(sys), out))
        return out

    def relax(self, step, max_iter, conv_crit, cg_tol=0.0, on_progress=None, poll_s=0.002):
        """Relax all systems.  ``on_progress(sys, rows)`` -- ``rows`` the (n, 3) relaxation record logged so far -- is
        called from THIS thread whenever the device has logged new rows (jf_relax runs in a helper thread meanwhile and
        keeps enqueueing Newton steps; ctypes releases the GIL), and once more with the complete record at the end."""
        relrec = np.zeros((self.n_sys, max_iter + 1, 3))
        cgit = np.zeros((self.n_sys, max(max_iter, 1)), np.int32)
        n_iter = np.zeros(self.n_sys, np.int32)
        call = lambda: self._L.jf_relax(self._h, float(step), int(max_iter), float(conv_crit), float(cg_tol), relrec, _ptr(cgit), n_iter)
        if on_progress is None:
            check(call())
        else:
            import threading
            import time
            box = {}

            def work():
                try:
                    box["rc"] = call()
                    box["err"] = self._L.jf_last_error().decode("utf-8", "replace")   # thread-local in the library
                except BaseException as e:      # noqa: BLE001 -- re-raised in the caller's thread
                    box["exc"] = e
            th = threading.Thread(target=work, daemon=True)
            seen = [0] * self.n_sys
            buf = np.zeros((max_iter + 1, 3))
            n = C.c_int32()
            th.start()
            while th.is_alive():
                th.join(poll_s)
                for s_ in range(self.n_sys):
                    if self._L.jf_relax_progress(self._h, s_, C.byref(n), _ptr(buf), max_iter + 1) == 0 and n.value > seen[s_] \
                            and th.is_alive():
                        seen[s_] = n.value
                        on_progress(s_, buf[: n.value].copy())
            if "exc" in box:
                raise box["exc"]
            if box["rc"] < 0:
                raise JFError("jf_b200 error %d: %s" % (box["rc"], box.get("err", "")))
            for s_ in range(self.n_sys):
                if n_iter[s_] + 1 > seen[s_]:
                    on_progress(s_, relrec[s_, : n_iter[s_] + 1].copy())
        return ([relrec[s, : n_iter[s] + 1].copy() for s in range(self.n_sys)],
                [cgit[s, : n_iter[s]].copy() for s in range(self.n_sys)], n_iter)

    def set_curve_bins(self, coords, r_inner, x_edges):
        """Classify the nodes into the radial bins [x_i, x_i+1) of |x|/r_inner with numpy's own comparisons
        (reference simulation.py:296-299) and hand the per-bin node lists to the device."""
        u = np.sqrt(np.sum(np.asarray(coords, dtype=np.float64) ** 2., axis=1)) / r_inner
        x_edges = np.asarray(x_edges, dtype=np.float64)
        lists = [np.nonzero((u >= x_edges[i]) & (u < x_edges[i + 1]))[0] for i in range(len(x_edges) - 1)]
        ptr = np.zeros(len(lists) + 1, np.int32)
        ptr[1:] = np.cumsum([len(l) for l in lists])
        nodes = np.ascontiguousarray(np.concatenate(lists) if ptr[-1] else np.zeros(0), dtype=np.int32)
        self._n_bins = len(lists)
        check(self._L.jf_set_curve_bins(self._h, len(lists), ptr, _ptr(nodes) if len(nodes) else None, float(r_inner)))

    def curves(self):
        out = np.empty((self.n_sys, self._n_bins))
        check(self._L.jf_get_curves(self._h, out))
        return out

    def stats(self):
        s = Stats()
        check(self._L.jf_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def reset_stats(self):
        check(self._L.jf_reset_stats(self._h))

    def info(self):
        i = Info()
        check(self._L.jf_get_info(self._h, C.byref(i)))
        return i.as_dict()


def solve_boundarycondition(nodes, tets, movable, f_target, U0, material, step, max_iter, conv_crit, beams=None, device=0):
    """One-call host-buffer solve (jf_solve_boundarycondition).  Returns dict(U, forces, relrec,
    cg_iters, n_iter, stats)."""
    L = load()
    nodes = np.ascontiguousarray(nodes, dtype=np.float64)
    tets = np.ascontiguousarray(tets, dtype=np.int32)
    movable = np.ascontiguousarray(movable, dtype=np.uint8)
    ft = np.ascontiguousarray(f_target, dtype=np.float64)
    U = np.array(U0, dtype=np.float64, order="C", copy=True)
    if beams is not None:
        beams = np.ascontiguousarray(beams, dtype=np.float64)
    relrec = np.zeros((max_iter + 1, 3))
    cgit = np.zeros(max(max_iter, 1), np.int32)
    n_iter = np.zeros(1, np.int32)
    forces = np.empty_like(nodes)
    st = Stats()
    k, d0, ls, ds = map(float, material)
    check(L.jf_solve_boundarycondition(int(device), len(nodes), nodes, len(tets), tets, movable, ft, U,
                                       0 if beams is None else len(beams), _ptr(beams), k, d0, ls, ds, float(step),
                                       int(max_iter), float(conv_crit), relrec, _ptr(cgit), n_iter, _ptr(forces), C.byref(st)))
    n = int(n_iter[0])
    return dict(U=U, forces=forces, relrec=relrec[: n + 1].copy(), cg_iters=cgit[:n].copy(), n_iter=n, stats=st.as_dict())


def bench_fp64(device=0):
    g = C.c_double()
    check(load().jf_bench_fp64(int(device), C.byref(g)))
    return g.value
