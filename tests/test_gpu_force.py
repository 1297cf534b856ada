"""Table interpolants, infer_pressure and reconstruct on the device against the reference's own functions.

The golden vectors in tests/golden/reference_lookup.npz were produced by the reference's create_lookup_functions
(simulation.py:595-621), infer_pressure (force.py:399-432) and reconstruct (force.py:13-210), executed in the build
container by tests/golden/make_reference_functions.py on the shipped Collagen12_BatchA table.
Tolerance: 1e-11 relative (libm vs CUDA log/exp/atan2, summation order of the means); counts, NaN patterns,
distances and medians' order statistics exact.
"""
import os

import numpy as np
import pytest

import jointforces_b200 as jf
from jointforces_b200 import force

pytestmark = pytest.mark.gpu
RTOL = 1e-11


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_lookup.npz"))


@pytest.fixture(scope="module")
def lookup(golden):
    table = {"pressure": golden["table_pressure"], "x": golden["table_x"], "y": golden["table_y"]}
    return force.create_lookup_functions(table)


def synthetic_series(n_frames=5, grid=40, seed=3):
    """Same construction as tests/golden/make_reference_functions.py:synthetic_series."""
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
        v[3, 5] = np.nan
        u[7, 9], v[7, 9] = 0.0, 0.0
        series.append(({'x': x, 'y': y, 'u': u, 'v': v}, {'mask': dist < r, 'radius': r, 'centroid': (cx, cy)}))
    return series


def assert_close(a, b, rtol=RTOL):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape
    assert np.array_equal(np.isnan(a), np.isnan(b)), "NaN pattern differs at %s" % np.flatnonzero(np.isnan(a) != np.isnan(b))[:5]
    ok = ~np.isnan(a)
    assert np.all(np.abs(a[ok] - b[ok]) <= rtol * np.abs(b[ok])), np.max(np.abs(a[ok] / b[ok] - 1))


def test_lookup_functions_match_reference(golden, lookup):
    get_displacement, get_pressure = lookup
    assert_close(get_displacement(golden["q_distance"], golden["q_pressure"]), golden["displacement"])
    assert_close(get_pressure(golden["q_distance"], golden["q_displacement"]), golden["pressure"])
    assert np.isnan(golden["pressure"]).sum() > 100 and (~np.isnan(golden["pressure"])).sum() > 3000     # both branches exercised
    # scalars and broadcasting like scipy's __call__
    d = get_displacement(2.0, 100.0)
    assert d.shape == () and abs(float(get_pressure(2.0, d)) / 100.0 - 1) < 1e-2     # two triangulations: inverse to table accuracy
    grid = get_displacement(np.array([[2.0], [4.0]]), np.array([10.0, 100.0, 1000.0]))
    assert grid.shape == (2, 3) and grid[0, 1] == d
    assert get_pressure(np.zeros(0), np.zeros(0)).shape == (0,)


@pytest.mark.parametrize("tag,angle_filter", [("f20", 20), ("f45", 45.5), ("none", None)])
def test_infer_pressure_matches_reference(golden, lookup, tag, angle_filter):
    series = synthetic_series()
    r0 = series[0][1]['radius']
    for f, (dis, seg) in enumerate(series):
        out = force.infer_pressure(np.ravel(dis['x']), np.ravel(dis['y']), np.ravel(dis['u']), np.ravel(dis['v']),
                                   seg['centroid'][0], seg['centroid'][1], r0, lookup[1], angle_filter=angle_filter)
        ref = [golden["infer_%d_%s_%s" % (f, tag, name)] for name in ("distance", "displacement", "angle", "pressure")]
        assert len(out) == 4 and all(len(o) == len(ref[0]) for o in out), (len(out[0]), len(ref[0]))
        assert np.array_equal(out[0], ref[0])                                   # IEEE sub/div/mul/add/sqrt: same bits
        # the projection is a two-term dot product (FMA or not inside BLAS): last-bit differences of its terms,
        # which unfiltered, nearly tangential vectors amplify in relative terms
        assert np.array_equal(np.isnan(out[1]), np.isnan(ref[1]))              # NaN v with a finite u: kept when unfiltered
        assert np.all(np.abs(out[1] - ref[1])[~np.isnan(ref[1])] <= 1e-15 * np.nanmax(np.abs(ref[1])))
        assert np.all(np.abs(out[2] - ref[2]) <= 1e-15)
        assert_close(out[3], ref[3], rtol=RTOL if angle_filter is not None else 1e-9)
        assert len(ref[0]) > 500


def test_infer_pressure_edge_cases(lookup):
    get_pressure = lookup[1]
    empty = force.infer_pressure([], [], [], [], 0.0, 0.0, 1.0, get_pressure)
    assert [len(e) for e in empty] == [0, 0, 0, 0]
    nan4 = force.infer_pressure([1.0, 2.0], [1.0, 2.0], [np.nan, np.nan], [0.0, 0.0], 0.0, 0.0, 1.0, get_pressure, angle_filter=None)
    assert [len(e) for e in nan4] == [0, 0, 0, 0]
    # a vector pointing away from the spheroid fails the filter; without the filter it is kept with a NaN pressure
    away = force.infer_pressure([300.0], [0.0], [5.0], [0.0], 0.0, 0.0, 100.0, get_pressure)
    assert len(away[0]) == 0
    kept = force.infer_pressure([300.0], [0.0], [5.0], [0.0], 0.0, 0.0, 100.0, get_pressure, angle_filter=None)
    assert kept[0][0] == 3.0 and kept[1][0] == -0.05 and kept[2][0] == 0.0 and np.isnan(kept[3][0])
    with pytest.raises(TypeError, match="device"):
        force.infer_pressure([1.0], [1.0], [1.0], [1.0], 0.0, 0.0, 1.0, lambda a, b: a)
    with pytest.raises(ValueError):
        force.infer_pressure([1.0, 2.0], [1.0], [1.0], [1.0], 0.0, 0.0, 1.0, get_pressure)


@pytest.mark.parametrize("tag,kw", [("default", {}), ("rmax", {"r_min": 1.5, "r_max": 6.0, "angle_filter": None}),
                                    ("continuous", {"continuous_radii": True, "r_min": None})])
def test_reconstruct_matches_reference(golden, tmp_path, tag, kw):
    series = synthetic_series()
    table = {"pressure": golden["table_pressure"], "x": golden["table_x"], "y": golden["table_y"]}
    jf.simulation.save_lookup_table(table, str(tmp_path / "table.npy"))
    for f, (dis, seg) in enumerate(series):
        np.save(tmp_path / ("def%06d.npy" % f), dis)
        np.save(tmp_path / ("seg%06d.npy" % f), seg)
    df = force.reconstruct(str(tmp_path), str(tmp_path / "table.npy"), 1.29, **kw)
    assert list(df.columns) == list(golden["reconstruct_columns"])
    ref = golden["reconstruct_" + tag]
    got = df.to_numpy(dtype=np.float64)
    assert got.shape == ref.shape == (5, 12)
    assert np.array_equal(got[:, 10:], ref[:, 10:])          # angles of the extreme sectors
    assert_close(got[:, :10], ref[:, :10])
    assert os.path.exists(tmp_path / "parameters_force.yml") and os.path.exists(tmp_path / "AppliedLookupTable.pkl")
    assert os.path.exists(tmp_path / "result.xlsx") or os.path.exists(tmp_path / "result.csv")


def test_not_accumulated_series_and_natural_order(golden, tmp_path):
    """Old standard: dis*.npy hold increments that reconstruct accumulates; files sort naturally (dis2 before dis10)."""
    series = synthetic_series(n_frames=3)
    table = {"pressure": golden["table_pressure"], "x": golden["table_x"], "y": golden["table_y"]}
    jf.simulation.save_lookup_table(table, str(tmp_path / "table.npy"))
    acc = tmp_path / "acc"
    inc = tmp_path / "inc"
    acc.mkdir(), inc.mkdir()
    names = ["8", "9", "10"]
    prev_u = prev_v = 0.0
    for name, (dis, seg) in zip(names, series):
        np.save(acc / ("def%s.npy" % name), dis)
        np.save(acc / ("seg%s.npy" % name), seg)
        np.save(inc / ("dis%s.npy" % name), {'x': dis['x'], 'y': dis['y'], 'u': dis['u'] - prev_u, 'v': dis['v'] - prev_v})
        np.save(inc / ("seg%s.npy" % name), seg)
        prev_u, prev_v = dis['u'], dis['v']
    a = force.reconstruct(str(acc), str(tmp_path / "table.npy"), 1.29, write=False).to_numpy(dtype=np.float64)
    b = force.reconstruct(str(inc), str(tmp_path / "table.npy"), 1.29, write=False).to_numpy(dtype=np.float64)
    assert a[0, 0] < a[1, 0] < a[2, 0]                        # frames in natural order: pressure grows with the frame
    assert_close(b, a, rtol=1e-6)                             # u - prev + prev is not exactly u


def test_statistics_against_numpy_on_a_large_frame(lookup):
    """100 k vectors in one frame: device mean / median / std / sector means against numpy on the device's own
    per-vector pressures (the reference's statistic code, force.py:95-121, restated)."""
    rng = np.random.default_rng(5)
    n = 100_000
    ang = rng.uniform(-np.pi, np.pi, n)
    dist = np.exp(rng.uniform(np.log(1.2), np.log(30.0), n))
    r = 50.0
    x, y = 1000.0 + r * dist * np.cos(ang), 1000.0 + r * dist * np.sin(ang)
    amp = r * 0.4 * dist ** -1.3 * np.exp(rng.normal(0, 0.2, n))
    u, v = -amp * np.cos(ang + rng.normal(0, 0.2, n)), -amp * np.sin(ang + rng.normal(0, 0.2, n))
    u[rng.random(n) < 0.05] = np.nan
    d, disp, a, p = force.infer_pressure(x, y, u, v, 1000.0, 1000.0, r, lookup[1], angle_filter=20)
    res = force.reconstruct_frames([(x, y, u, v, 1000.0, 1000.0, r)], lookup[1], r_min=2, angle_filter=20)
    m = d > 2
    assert int(res['n_used'][0]) == int((~np.isnan(p[m])).sum()) > 30000
    assert abs(res['pressure_mean'][0] / np.nanmean(p[m]) - 1) < 1e-12
    assert res['pressure_median'][0] == np.nanmedian(p[m])                       # exact order statistic
    assert abs(res['pressure_std'][0] / np.nanstd(p[m], ddof=1) - 1) < 1e-11
    with np.errstate(all="ignore"):
        sect = [np.nanmean(p[m & (a >= (al - 5) * np.pi / 180.) & (a < (al + 5) * np.pi / 180.)]) for al in range(-180, 180, 5)]
    assert_close(res['pressure_angles'][0], np.array(sect), rtol=1e-12)
    # an even count takes the mean of the two middle values; all-NaN frames give NaN everywhere
    res2 = force.reconstruct_frames([(x[:7], y[:7], u[:7], v[:7], 1000.0, 1000.0, r), (x[:3], y[:3], np.full(3, np.nan), v[:3], 0, 0, r)],
                                    lookup[1], r_min=None, angle_filter=None)
    d7, _, _, p7 = force.infer_pressure(x[:7], y[:7], u[:7], v[:7], 1000.0, 1000.0, r, lookup[1], angle_filter=None)
    with np.errstate(all="ignore"):
        assert res2['pressure_median'][0] == np.nanmedian(p7) or (np.isnan(res2['pressure_median'][0]) and np.all(np.isnan(p7)))
    assert res2['n_used'][1] == 0 and np.isnan(res2['pressure_mean'][1]) and np.all(np.isnan(res2['pressure_angles'][1]))


def test_random_scattered_table_against_scipy():
    """Not a regular table: random points, values with a gradient; scipy's LinearNDInterpolator is the reference's
    dependency for this step and available wherever the tests run."""
    from scipy.interpolate import LinearNDInterpolator
    rng = np.random.default_rng(9)
    pts = rng.normal(size=(3000, 2)) * [2.0, 0.5]
    vals = np.sin(pts[:, 0]) + pts[:, 1] ** 2
    q = rng.normal(size=(50_000, 2)) * [2.5, 0.7]
    ref = LinearNDInterpolator(pts, vals)(q[:, 0], q[:, 1])
    t = force.TableInterpolant(pts, vals)
    got = t(q[:, 0], q[:, 1])
    same = np.isnan(got) == np.isnan(ref)
    assert (~same).sum() <= 2, (~same).sum()                   # hull edge, decided by the last bit
    ok = same & ~np.isnan(ref)
    assert ok.sum() > 40000
    assert np.max(np.abs(got[ok] - ref[ok])) < 1e-12
    assert np.array_equal(t(pts[:100, 0], pts[:100, 1]), vals[:100]) or np.max(np.abs(t(pts[:100, 0], pts[:100, 1]) - vals[:100])) < 1e-13
    t.close()


def test_table_error_paths():
    """Bad triangulations and arguments are refused with an error code and a message, never a crash."""
    import ctypes as C
    from jointforces_b200 import _lib
    L = _lib.load()
    pts = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [1.0, 1.0]])
    vals = np.arange(4.0)
    from scipy.spatial import Delaunay
    tri = Delaunay(pts)
    simp = np.ascontiguousarray(tri.simplices, dtype=np.int32)
    trans = np.ascontiguousarray(tri.transform, dtype=np.float64)
    nbr = np.ascontiguousarray(tri.neighbors, dtype=np.int32)
    h = _lib._vp()
    bad = simp.copy()
    bad[0, 0] = 7                                                     # vertex outside the point set
    assert L.jf_table_create(C.byref(h), 0, 4, pts, vals, len(simp), bad, trans, nbr) == -1
    assert b"outside the point set" in L.jf_last_error()
    assert L.jf_table_create(C.byref(h), 0, 2, pts[:2], vals[:2], len(simp), simp, trans, nbr) == -1     # fewer than 3 points
    assert L.jf_table_create(C.byref(h), 99, 4, pts, vals, len(simp), simp, trans, nbr) == -1            # no such device
    same_x = np.array([[0.0, 0.0], [0.0, 1.0], [0.0, 2.0]])
    assert L.jf_table_create(C.byref(h), 0, 3, same_x, vals[:3], 1, np.zeros((1, 3), np.int32) + np.arange(3, dtype=np.int32),
                             np.zeros((1, 3, 2)), -np.ones((1, 3), np.int32)) == -1
    assert b"abscissa" in L.jf_last_error()
    # a valid table: corners reproduce the values, the centre is their mean, outside is NaN
    t = force.TableInterpolant(pts, vals)
    assert np.array_equal(t(pts[:, 0], pts[:, 1]), vals)
    assert abs(float(t(0.5, 0.5)) - 1.5) < 1e-15 and np.isnan(t(1.5, 0.5)) and np.isnan(t(np.nan, 0.5))
    with pytest.raises(ValueError):
        force.TableInterpolant(np.zeros((3, 3)), np.zeros(3))
    t.close()
    t.close()                                                          # idempotent
