"""Force reconstruction: the consumer of the lookup table, on the device.

Mirror of the data path of the reference's ``jointforces/force.py`` and of the table interpolants it loads
(``jointforces/simulation.py:595-621,643-660``):

``create_lookup_functions`` :595, ``load_lookup_functions`` :643  ->  ``(get_displacement, get_pressure)`` with the
reference's signatures, evaluated by ``jf_table_eval`` on the GPU;
``infer_pressure`` force.py:399-432 (same signature, same four outputs) -> one fused kernel, ``jf_infer_pressure``;
``reconstruct`` force.py:13-210 -> all frames of a series in one ``jf_reconstruct_frames`` call (``reconstruct_frames``
is the in-memory form); the DataFrame has the reference's twelve columns and file names.

The triangulation behind the interpolants is the one scipy's ``LinearNDInterpolator`` builds
(``scipy.spatial.Delaunay(points)``, same Qhull, same options) -- made once per table on the host; every evaluation
runs on the device and raises :class:`jointforces_b200._lib.JFError` without one (no CPU fallback).
"""
from __future__ import annotations

import ctypes as C
import os
import re
from glob import glob

import numpy as np

from . import _lib

N_SECTORS = 72
FRAME_OUT = 4 + N_SECTORS


class TableInterpolant:
    """One ``LinearNDInterpolator(points, values)`` living on the GPU."""

    def __init__(self, points, values, device=0):
        from scipy.spatial import Delaunay
        points = np.ascontiguousarray(points, dtype=np.float64)
        values = np.ascontiguousarray(values, dtype=np.float64)
        if points.ndim != 2 or points.shape[1] != 2 or values.shape != (len(points),):
            raise ValueError("points must be (n, 2) and values (n,)")
        tri = Delaunay(points)          # what LinearNDInterpolator does internally
        self.n_points, self.n_simplices = len(points), len(tri.simplices)
        L = _lib.load()
        self._L = L
        self._h = _lib._vp()
        _lib.check(L.jf_table_create(C.byref(self._h), int(device), len(points), points, values, len(tri.simplices),
                                     np.ascontiguousarray(tri.simplices, dtype=np.int32),
                                     np.ascontiguousarray(tri.transform, dtype=np.float64),
                                     np.ascontiguousarray(tri.neighbors, dtype=np.int32)))

    def close(self):
        if getattr(self, "_h", None):
            self._L.jf_table_destroy(self._h)
            self._h = None

    __del__ = close

    def __call__(self, x, y, log_x=False, log_y=False, exp_out=False):
        """f(x, y) for broadcastable x, y (scalars or arrays), NaN outside the hull, like scipy's ``__call__``."""
        x, y = np.broadcast_arrays(np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64))
        shape = x.shape
        xf, yf = np.ascontiguousarray(x.ravel()), np.ascontiguousarray(y.ravel())
        out = np.empty(xf.size, dtype=np.float64)
        _lib.check(self._L.jf_table_eval(self._h, xf.size, xf, yf, int(log_x), int(log_y), int(exp_out), out))
        return out.reshape(shape)

    @property
    def launches(self):
        return int(self._L.jf_table_launches(self._h))


class _GetDisplacement:
    def __init__(self, table):
        self.table = table

    def __call__(self, distance, pressure):
        return self.table(distance, pressure, log_x=True, log_y=True)


class _GetPressure:
    def __init__(self, table):
        self.table = table

    def __call__(self, distance, displacement):
        return self.table(distance, displacement, log_x=True, exp_out=True)


def create_lookup_functions(lookup_table, device=0):
    """``get_displacement(distance, pressure)`` and ``get_pressure(distance, displacement)`` of a lookup table
    ``{'pressure': (P,), 'x': (X,), 'y': (P, X)}`` -- reference simulation.py:595-621, evaluated on the GPU."""
    pressure = np.asarray(lookup_table['pressure'])
    distance = np.asarray(lookup_table['x'])
    displacement = np.asarray(lookup_table['y'])

    x, y = np.meshgrid(np.log(distance), np.log(pressure))
    mask = ~(np.isnan(x) | np.isnan(y) | np.isnan(displacement))
    x, y, displacement = x[mask], y[mask], displacement[mask]

    forward = TableInterpolant(np.array([x, y]).T, displacement, device)
    inverse = TableInterpolant(np.array([x, displacement]).T, y, device)
    return _GetDisplacement(forward), _GetPressure(inverse)


def load_lookup_functions(file, device=0):
    """Reference simulation.py:643-660 for the current format (a pickled dict written by ``save_lookup_table``)."""
    lookup_table = np.load(file, allow_pickle=True).item()
    get_displacement, get_pressure = create_lookup_functions(lookup_table, device)
    print("Found Lookuptable, and created lookup functions")
    return get_displacement, get_pressure


def _table_of(get_pressure):
    if not isinstance(get_pressure, _GetPressure):
        raise TypeError("get_pressure must come from jointforces_b200.force.create_lookup_functions / load_lookup_functions "
                        "(the inference runs on the device; there is no host path for arbitrary callables)")
    return get_pressure.table


def _filter_args(angle_filter):
    if angle_filter is None:
        return 0, 0.0
    return 1, float(np.cos(2 * np.pi * angle_filter / 360))


def infer_pressure(x_rav, y_rav, u_rav, v_rav, x_sph, y_sph, r_sph, get_pressure, angle_filter=20):
    """Reference force.py:399-432: relative distance, displacement projected on the spheroid axis, angle and the
    pressure the table assigns, for the vectors with a non-NaN ``u`` that pass the angle filter (``None``: all)."""
    t = _table_of(get_pressure)
    arrs = [np.ascontiguousarray(np.ravel(a), dtype=np.float64) for a in (x_rav, y_rav, u_rav, v_rav)]
    n = arrs[0].size
    if any(a.size != n for a in arrs):
        raise ValueError("x, y, u, v must have the same number of elements")
    out = [np.empty(n, dtype=np.float64) for _ in range(4)]
    use, cos_min = _filter_args(angle_filter)
    n_out = C.c_int64()
    _lib.check(t._L.jf_infer_pressure(t._h, n, *arrs, float(x_sph), float(y_sph), float(r_sph), use, cos_min, *out, C.byref(n_out)))
    return [o[:n_out.value] for o in out]


SECTOR_ANGLES = list(range(-180, 180, 5))


def reconstruct_frames(frames, get_pressure, r_min=2, angle_filter=20, r_max=None):
    """The per-frame loop of ``reconstruct`` (force.py:70-121) for a whole series in one device call.

    ``frames``: sequence of ``(x, y, u, v, cx, cy, r)`` -- raveled PIV grid and accumulated deformation of a frame,
    centroid and the radius to normalise with.  Returns a dict with ``pressure_mean/median/std`` (n_frames,),
    ``n_used`` and ``pressure_angles`` (n_frames, 72) for alpha = -180, -175 .. 175."""
    t = _table_of(get_pressure)
    if r_max and r_min is None:
        raise TypeError("'>' not supported between instances of 'float' and 'NoneType'")   # what the reference's mask raises
    n_frames = len(frames)
    if n_frames == 0:
        return {'pressure_mean': np.zeros(0), 'pressure_median': np.zeros(0), 'pressure_std': np.zeros(0),
                'n_used': np.zeros(0, dtype=np.int64), 'pressure_angles': np.zeros((0, N_SECTORS))}
    parts = [[np.ravel(np.asarray(f[k], dtype=np.float64)) for f in frames] for k in range(4)]
    sizes = [p.size for p in parts[0]]
    for k in range(1, 4):
        if [p.size for p in parts[k]] != sizes:
            raise ValueError("x, y, u, v of a frame must have the same number of elements")
    frame_ptr = np.zeros(n_frames + 1, dtype=np.int64)
    np.cumsum(sizes, out=frame_ptr[1:])
    x, y, u, v = [np.ascontiguousarray(np.concatenate(p)) if frame_ptr[-1] else np.zeros(1) for p in parts]
    geom = np.ascontiguousarray([[float(f[4]), float(f[5]), float(f[6])] for f in frames], dtype=np.float64)
    bounds = np.ascontiguousarray([[(a - 5) * np.pi / 180., (a + 5) * np.pi / 180.] for a in SECTOR_ANGLES], dtype=np.float64)
    stats = np.empty((n_frames, FRAME_OUT), dtype=np.float64)
    use, cos_min = _filter_args(angle_filter)
    # `if r_max:` in the reference: 0 and None both mean "no upper limit"
    _lib.check(t._L.jf_reconstruct_frames(t._h, n_frames, frame_ptr, geom, x, y, u, v, use, cos_min,
                                          float('nan') if r_min is None else float(r_min),
                                          float(r_max) if r_max else float('nan'), bounds, stats))
    return {'pressure_mean': stats[:, 0].copy(), 'pressure_median': stats[:, 1].copy(), 'pressure_std': stats[:, 2].copy(),
            'n_used': stats[:, 3].astype(np.int64), 'pressure_angles': stats[:, 4:].copy()}


def _natural_key(name):
    return [int(tok) if tok.isdigit() else tok.lower() for tok in re.split(r'(\d+)', name)]


def _load(file):
    return np.load(file, allow_pickle=True).item()      # reference utils.py:4-5


COLUMNS = ['Mean Pressure (Pa)', 'Median Pressure (Pa)', 'St.dev. Pressure (Pa)',
           'Mean Contractility (µN)', 'Median Contractility (µN)', 'St.dev. Contractility (µN)',
           'Minimal Median Pressure (Pa)', 'Maximal Median Pressure (Pa)',
           'Minimal Median Contractility (µN)', 'Maximal Median Contractility (µN)',
           'Angle of minimal Pr./Contr. (deg)', 'Angle of maximal Pr./Contr. (deg)']


def reconstruct(folder, lookupfile, muperpixel, outfile=None, r_min=2, angle_filter=20, r_max=None, continuous_radii=False,
                write=True):
    """Reference force.py:13-210: contractility over time from the PIV results in ``folder`` (``seg*.npy`` and
    ``def*.npy`` -- or, old standard, ``dis*.npy`` that still have to be accumulated) and a lookup table.

    Returns the reference's DataFrame (twelve columns, one row per frame); with ``write`` it also stores
    ``result.xlsx`` / ``result_angles_pressure.xlsx`` / ``result_angles_contractility.xlsx`` (or ``outfile`` and its
    two companions) like the reference -- as ``.csv`` next to those names when pandas has no Excel writer --, a copy of
    the table and ``parameters_force.yml``."""
    import pandas as pd
    seg_files = sorted(glob(folder + '/seg*.npy'), key=_natural_key)
    d_accumulated_files = sorted(glob(folder + '/def*.npy'), key=_natural_key)
    d_notaccumulated_files = sorted(glob(folder + '/dis*.npy'), key=_natural_key)
    if len(d_notaccumulated_files) > len(d_accumulated_files):
        accumulated = False
        dis_files = d_notaccumulated_files
        print("Found not-accumulated deformation files (old standard) and will conduct calculations accordingly.")
    else:
        accumulated = True
        dis_files = d_accumulated_files

    get_displacement, get_pressure = load_lookup_functions(lookupfile)

    r0 = _load(seg_files[0])['radius']
    A0 = 4 * np.pi * (r0 * (10 ** -6)) ** 2

    frames = []
    u_sum = v_sum = None
    for dis_file, seg_file in zip(dis_files, seg_files):
        dis, seg = _load(dis_file), _load(seg_file)
        if accumulated or u_sum is None:
            u_sum, v_sum = np.ravel(dis['u']).astype(np.float64), np.ravel(dis['v']).astype(np.float64)
        else:
            u_sum = u_sum + np.ravel(dis['u'])
            v_sum = v_sum + np.ravel(dis['v'])
        cx, cy = seg['centroid']
        if continuous_radii:
            r0 = seg['radius']
        frames.append((np.ravel(dis['x']), np.ravel(dis['y']), u_sum, v_sum, cx, cy, r0))

    res = reconstruct_frames(frames, get_pressure, r_min=r_min, angle_filter=angle_filter, r_max=r_max)
    scale = A0 * (muperpixel ** 2.) * (10 ** 6)     # pressure -> contractility in µN
    pr_angles = res['pressure_angles']
    n = len(frames)
    p_min, p_max, a_min, a_max = (np.full(n, np.nan) for _ in range(4))
    for i in range(n):
        if not np.all(np.isnan(pr_angles[i])):       # nanargmin raises on an all-NaN row: the reference then stores NaN
            i_min, i_max = np.nanargmin(pr_angles[i]), np.nanargmax(pr_angles[i])
            p_min[i], p_max[i] = pr_angles[i, i_min], pr_angles[i, i_max]
            a_min[i], a_max[i] = SECTOR_ANGLES[i_min], SECTOR_ANGLES[i_max]
    df = pd.DataFrame({'pressure_mean': res['pressure_mean'], 'pressure_median': res['pressure_median'],
                       'pressure_std': res['pressure_std'],
                       'contractility_mean': res['pressure_mean'] * scale, 'contractility_median': res['pressure_median'] * scale,
                       'contractility_std': res['pressure_std'] * scale,
                       'pressure_min': p_min, 'pressure_max': p_max,
                       'contractility_min': p_min * scale, 'contractility_max': p_max * scale,
                       'angle_min': a_min, 'angle_max': a_max})
    df.columns = COLUMNS
    an = pd.DataFrame(pr_angles, columns=SECTOR_ANGLES)
    an2 = pd.DataFrame(pr_angles * scale, columns=SECTOR_ANGLES)

    if write:
        def to_excel(frame, path):
            try:
                frame.to_excel(path)
            except (ImportError, ModuleNotFoundError):
                frame.to_csv(os.path.splitext(path)[0] + '.csv')

        if outfile is not None:
            to_excel(df, outfile)
            to_excel(an, outfile[:-5] + '_angles_pressure.xlsx')
            to_excel(an2, outfile[:-5] + '_angles_contractility.xlsx')
        else:
            to_excel(df, folder + '//result.xlsx')
            to_excel(an, folder + '//result_angles_pressure.xlsx')
            to_excel(an2, folder + '//result_angles_contractility.xlsx')
        from shutil import copyfile
        copyname = "AppliedLookupTable.pkl"
        copyfile(lookupfile, os.path.join(folder, copyname))
        import yaml
        dict_file = {'Force': {'folder': [folder], 'lookupfile': [lookupfile], 'muperpixel': [muperpixel],
                               'r_min': [str(r_min)], 'r_max': [str(r_max)], 'angle_filter': [str(angle_filter)],
                               'continuous_radii': [continuous_radii], 'accumulated': [str(accumulated)],
                               'Copy of applied Lookuptable': [copyname]}}
        with open(os.path.join(folder, 'parameters_force.yml'), 'w') as yaml_file:
            yaml.dump(dict_file, yaml_file, default_flow_style=False)
    return df
