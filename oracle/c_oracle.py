"""ctypes front for oracle/saeno_oracle.c (TEST INFRASTRUCTURE ONLY; PARITY UNPINNED -- see the
header of saeno_oracle.c).  Builds the shared object on first use if it is missing."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsaeno_oracle.so")
_lib = None

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_lp = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_bp = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def build(force=False):
    src = os.path.join(_HERE, "saeno_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.so_material_eval.argtypes = [C.c_double] * 4 + [C.c_int, _dp, _dp, _dp, _dp]
        L.so_material_eval.restype = None
        L.so_shape_tensors.argtypes = [C.c_int, _dp, _ip, _dp, _dp]
        L.so_shape_tensors.restype = None
        L.so_element_eval.argtypes = [C.c_int, _ip, _dp, _dp, _dp, C.c_int, _dp] + [C.c_double] * 4 + [_dp, _dp, _dp]
        L.so_element_eval.restype = None
        L.so_relax.argtypes = [C.c_int, C.c_int, _dp, _ip, _bp, _dp, _dp, C.c_int, _dp] + [C.c_double] * 4 + \
                              [C.c_double, C.c_int, C.c_double, _dp, _ip, _dp, _dp, C.c_int, C.c_double]
        L.so_relax.restype = C.c_int
        L.so_cg_variants.argtypes = [C.c_int, C.c_int, _dp, _ip, _bp, _dp, _dp, C.c_int, _dp] + [C.c_double] * 5 + [_ip]
        L.so_cg_variants.restype = None
        L.so_pattern_nnz.argtypes = [C.c_int, C.c_int, _ip, _bp]
        L.so_pattern_nnz.restype = C.c_int64
        L.so_assemble.argtypes = [C.c_int, C.c_int, _dp, _ip, _bp, _dp, C.c_int, _dp] + [C.c_double] * 4 + \
                                 [_dp, _lp, _ip, _dp]
        L.so_assemble.restype = C.c_double
        _lib = L
    return _lib


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def material_eval(k, d0, ls, ds, lam):
    lam = _c(lam, np.float64).ravel()
    w0, w1, w2 = (np.empty_like(lam) for _ in range(3))
    lib().so_material_eval(k, d0, ls, ds, len(lam), lam, w0, w1, w2)
    return w0, w1, w2


def shape_tensors(nodes, tets):
    nodes, tets = _c(nodes, np.float64), _c(tets, np.int32)
    T = len(tets)
    phi, vol = np.empty((T, 4, 3)), np.empty(T)
    lib().so_shape_tensors(T, nodes, tets, phi, vol)
    return phi, vol


def element_eval(nodes, tets, U, beams, material):
    """-> E (T,), f_el (T,4,3), K_el (T,4,4,3,3).  material = (k, d0, ls, ds)."""
    nodes, tets, U, beams = _c(nodes, np.float64), _c(tets, np.int32), _c(U, np.float64), _c(beams, np.float64)
    phi, vol = shape_tensors(nodes, tets)
    T = len(tets)
    E, f, K = np.empty(T), np.empty((T, 4, 3)), np.empty((T, 4, 4, 3, 3))
    lib().so_element_eval(T, tets, U, phi, vol, len(beams), beams, *map(float, material), E, f, K)
    return E, f, K


def assemble(nodes, tets, movable, U, beams, material):
    """-> strain energy, forces (N,3), scipy CSR (3N x 3N, rows of free nodes only)."""
    import scipy.sparse as ssp
    nodes, tets, U, beams = _c(nodes, np.float64), _c(tets, np.int32), _c(U, np.float64), _c(beams, np.float64)
    mov = _c(movable, np.uint8)
    N, T = len(nodes), len(tets)
    nnz = lib().so_pattern_nnz(N, T, tets, mov)
    forces = np.empty((N, 3))
    rowptr, col, val = np.empty(3 * N + 1, np.int64), np.empty(nnz, np.int32), np.empty(nnz)
    e = lib().so_assemble(N, T, nodes, tets, mov, U, len(beams), beams, *map(float, material), forces, rowptr, col, val)
    return e, forces, ssp.csr_matrix((val, col, rowptr), shape=(3 * N, 3 * N))


def relax(nodes, tets, movable, f_target, U0, beams, material, step, max_iter, conv_crit, nthreads=0, cg_tol=0.0):
    """Run the whole relaxation.  Returns dict(U, forces, relrec (n+1,3), cg_iters (n,), n_iter,
    t_element, t_cg)."""
    nodes, tets, beams = _c(nodes, np.float64), _c(tets, np.int32), _c(beams, np.float64)
    mov = _c(movable, np.uint8)
    ft = _c(np.nan_to_num(np.asarray(f_target, dtype=np.float64)), np.float64)
    U = _c(U0, np.float64).copy()
    N, T = len(nodes), len(tets)
    relrec = np.zeros((max_iter + 1, 3))
    cgit = np.zeros(max_iter, np.int32)
    forces = np.empty((N, 3))
    timing = np.zeros(2)
    n = lib().so_relax(N, T, nodes, tets, mov, ft, U, len(beams), beams, *map(float, material), float(step),
                       int(max_iter), float(conv_crit), relrec, cgit, forces, timing, int(nthreads), float(cg_tol))
    return dict(U=U, forces=forces, relrec=relrec[: n + 1].copy(), cg_iters=cgit[:n].copy(), n_iter=n,
                t_element=timing[0], t_cg=timing[1])


def cg_variants(nodes, tets, movable, f_target, U, beams, material, tol=1e-5):
    """CG iteration counts of one Newton step at displacements U under three arithmetic variants (plain double as the
    oracle, fused multiply-add + pairwise sums as a GPU, long double): tools/cg_precision_study.py."""
    nodes, tets, beams, U = _c(nodes, np.float64), _c(tets, np.int32), _c(beams, np.float64), _c(U, np.float64)
    ft = _c(np.nan_to_num(np.asarray(f_target, dtype=np.float64)), np.float64)
    it = np.zeros(3, np.int32)
    lib().so_cg_variants(len(nodes), len(tets), nodes, tets, _c(movable, np.uint8), ft, U, len(beams), beams,
                         *map(float, material), float(tol), it)
    return dict(double_plain=int(it[0]), double_fma_pairwise=int(it[1]), long_double=int(it[2]))
