"""Developer timing of the table consumer: reconstruct_frames on the device against the reference's per-frame
numpy/scipy loop (restated from force.py:70-121, 399-432) on the host.  python tools/force_bench.py [FRAMES] [GRID]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.interpolate import LinearNDInterpolator

from jointforces_b200 import force

n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 100
grid = int(sys.argv[2]) if len(sys.argv) > 2 else 128
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "reference_lookup.npz"))
table = {"pressure": g["table_pressure"], "x": g["table_x"], "y": g["table_y"]}

rng = np.random.default_rng(0)
xs = np.arange(16, 16 + 16 * grid, 16, dtype=np.float64)
x, y = np.meshgrid(xs, xs)
c = xs.mean()
frames = []
for f in range(n_frames):
    r = 90.0
    dx, dy = x - c, y - c
    dist = np.sqrt(dx ** 2 + dy ** 2)
    amp = 0.05 * (f + 1) * r * (r / np.maximum(dist, r)) ** 1.2
    u = -amp * dx / dist + rng.normal(0, 0.1, x.shape)
    v = -amp * dy / dist + rng.normal(0, 0.1, x.shape)
    u[dist < r] = np.nan
    frames.append((x.ravel(), y.ravel(), u.ravel(), v.ravel(), c, c, r))

t0 = time.time()
gd, gp = force.create_lookup_functions(table)
t_setup = time.time() - t0
res = force.reconstruct_frames(frames, gp)          # warm-up
t0 = time.time()
for _ in range(3):
    res = force.reconstruct_frames(frames, gp)
t_dev = (time.time() - t0) / 3

# host loop of the reference
t0 = time.time()
lx, ly = np.meshgrid(np.log(table['x']), np.log(table['pressure']))
ok = ~np.isnan(table['y'])
f_inv = LinearNDInterpolator(np.array([lx[ok], table['y'][ok]]).T, ly[ok])
t_host_setup = time.time() - t0
t0 = time.time()
means = []
with np.errstate(all="ignore"):
    for (xr, yr, ur, vr, cx, cy, r) in frames:
        sel = ~np.isnan(ur)
        dx, dy = (xr[sel] - cx) / r, (yr[sel] - cy) / r
        distance = np.sqrt(dx ** 2 + dy ** 2)
        angle = np.arctan2(dy, dx)
        u2, v2 = ur[sel] / r, vr[sel] / r
        m = np.array([dx, dy]).T / distance[:, None]
        d = np.array([u2, v2]).T
        displacement = np.array([-np.dot(di, mi) for di, mi in zip(d, m)])
        mask = displacement / np.sqrt(u2 ** 2. + v2 ** 2.) > np.cos(2 * np.pi * 20 / 360)
        distance, angle, displacement = distance[mask], angle[mask], displacement[mask]
        pressure = np.exp(f_inv(np.log(distance), displacement))
        mk = distance > 2
        for alpha in range(-180, 180, 5):
            mask2 = (angle >= (alpha - 5) * np.pi / 180.) & (angle < (alpha + 5) * np.pi / 180.)
            np.nanmean(pressure[mk & mask2])
        means.append((np.nanmean(pressure[mk]), np.nanmedian(pressure[mk]), np.nanstd(pressure[mk], ddof=1)))
t_host = time.time() - t0
means = np.array(means)
err = max(np.nanmax(np.abs(res['pressure_mean'] / means[:, 0] - 1)), np.nanmax(np.abs(res['pressure_std'] / means[:, 2] - 1)))
print("frames %d x %d vectors: device %.1f ms (%.0f frames/s, table setup %.0f ms) | host loop %.0f ms (%.1f frames/s, setup %.0f ms) "
      "| speed-up %.0fx | max rel diff mean/std %.1e, median %.1e" % (
          n_frames, grid * grid, t_dev * 1e3, n_frames / t_dev, t_setup * 1e3, t_host * 1e3, n_frames / t_host, t_host_setup * 1e3,
          t_host / t_dev, err, np.nanmax(np.abs(res['pressure_median'] / means[:, 1] - 1))))
