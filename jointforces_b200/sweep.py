"""Pressure sweep: many load cases on one mesh, spread over GPUs.

Replaces the process-per-pressure loop of ``/root/reference/jointforces/simulation.py:204-246``.
The solves are independent (every one is a cold start), so ranks share
nothing on the data path; the only collective is the final gather of the (n, 100) displacement
curves over NCCL.  One process per GPU: run under ``torchrun``/``torch.distributed.run`` for
several GPUs (RANK / WORLD_SIZE / LOCAL_RANK are read from the environment), or plainly for one.

Work distribution.  The cost of a solve varies by a factor of two over the pressure range (Newton
and CG iteration counts; ``pressure_cost``), and nothing bounds it in advance for a new mesh or
material, so with an initialised process group the ranks draw load cases from a shared counter in
the group's key-value store (one ``add`` per draw, most expensive cases first): whoever is free
takes the next case, the tail imbalance is at most one of the cheapest solves.  Without a store
(``schedule='lpt'``) the cases are dealt by longest-processing-time-first on the cost model.
Folder output (``parameters.txt``, ``relrec.txt``, ``solver.saenopy``) is written by a background
thread while the GPU relaxes the next case.
"""
from __future__ import annotations

import os
import warnings
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import _lib

MAX_BATCH = 16  # systems relaxed in lockstep per context (library limit is 64)
_SWEEP_CALLS = 0  # sweeps started by this process: names the shared counter of each sweep (same sequence on all ranks)


def _env_rank_world():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def _group_store():
    """The key-value store of the default process group (None without one)."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist.distributed_c10d._get_default_store()
    except Exception:
        pass
    return None


# Relative cost of one relaxation against log10(pressure / Pa): device milliseconds of the 150 solves of BASELINE
# configs[2] (collagen12, 17,612 nodes) on one B200, at every 15th pressure, normalised to the cheapest.  Used to
# ORDER the work (most expensive first) and for the static LPT schedule; the dynamic schedule does not depend on its
# accuracy.  Other materials shift the curve along the pressure axis by their stiffness ratio (K_0 / 1645).
_COST_LOGP = np.array([-1.0, -0.5, 0.0, 0.5, 1.0, 1.5, 2.0, 2.5, 3.0, 3.5, 4.0])
_COST_REL = np.array([1.00, 1.00, 1.02, 1.10, 1.25, 1.45, 1.60, 1.55, 1.40, 1.30, 1.25])


def pressure_cost(values, material=None):
    """Estimated relative cost of relaxing each pressure (see ``_COST_REL``)."""
    values = np.asarray(values, dtype=np.float64)
    k0 = float(material['K_0']) if material is not None else 1645.0
    logp = np.log10(np.maximum(values, 1e-300) * (1645.0 / k0))
    return np.interp(logp, _COST_LOGP, _COST_REL)


def lpt_assignment(costs, world):
    """Longest-processing-time-first: cases in descending cost, each to the least loaded rank (ties: lowest rank).
    Returns a list of index arrays (ascending) per rank.  Deterministic, so every rank computes the same deal."""
    costs = np.asarray(costs, dtype=np.float64)
    order = np.argsort(-costs, kind="stable")
    load = np.zeros(world)
    mine = [[] for _ in range(world)]
    for i in order:
        r = int(np.argmin(load))
        mine[r].append(int(i))
        load[r] += costs[i]
    return [np.sort(np.asarray(m, dtype=np.int64)) for m in mine]


def shard_indices(values, rank, world, material=None):
    """This rank's share of ``values`` under the static LPT schedule (ascending indices)."""
    return lpt_assignment(pressure_cost(values, material), world)[rank]


def work_order(values, costs, batch=1):
    """The order in which load cases are handed out: most expensive first.  With ``batch`` cases relaxed side by side
    (lockstep Newton steps, one CG launch for the batch) the members of a batch should follow similar trajectories, or
    the quicker one idles: batches are runs of ``batch`` NEIGHBOURS ON THE VALUE AXIS (similar pressure -> similar Newton
    and CG iteration counts; two pressures of equal cost on opposite sides of the cost peak differ by a factor of two in
    Newton steps), and the batches are ordered by their summed cost, dearest first."""
    costs = np.asarray(costs, dtype=np.float64)
    n = len(costs)
    if batch <= 1 or n <= batch:
        return np.argsort(-costs, kind="stable")
    by_value = np.argsort(np.asarray(values, dtype=np.float64), kind="stable")
    groups = [by_value[a:a + batch] for a in range(0, n, batch)]
    groups.sort(key=lambda g: -float(costs[g].sum() / len(g)))
    # a short last group (n not a multiple of batch) goes to the end: every draw before it is a full batch
    full = [g for g in groups if len(g) == batch]
    short = [g for g in groups if len(g) < batch]
    return np.concatenate(full + short)


class _Claimer:
    """Hands out positions 0..n-1 of the cost-ordered work list.  With a store: a shared counter (dynamic schedule);
    without: this rank's LPT share."""

    def __init__(self, order, costs, rank, world, schedule, batch=1):
        global _SWEEP_CALLS
        _SWEEP_CALLS += 1
        self.order = order
        self.n = len(order)
        self.world = world
        self.store = _group_store() if schedule == "dynamic" else None
        if self.store is not None:
            self.key = "jf_b200_sweep_%d" % _SWEEP_CALLS
            self.static = None
        elif world > 1:
            # static deal of whole batches (runs of `batch` entries of the order stay together on one rank)
            chunks = [order[a:a + batch] for a in range(0, self.n, max(1, batch))]
            chunk_cost = [float(np.sum(np.asarray(costs)[np.asarray(ch, dtype=np.int64)])) for ch in chunks]
            mine = lpt_assignment(chunk_cost, world)[rank]
            self.static = [int(i) for k in mine for i in chunks[int(k)]]
            self.pos = 0
        else:
            self.static = [int(i) for i in order]
            self.pos = 0

    def take(self, count):
        """Up to ``count`` load-case indices, [] when the list is exhausted.  When fewer cases are left in a shared list than
        there are ranks the draws shrink to single cases (measured on 8 GPUs: shrinking earlier, at 2 x ranks, costs 3 %:
        two rounds of single solves instead of one round of pairs)."""
        if self.store is not None:
            if count > 1 and self.world > 1:
                taken = int(self.store.add(self.key, 0))
                if self.n - taken <= self.world:   # fewer cases left than ranks: one each (a pair costs 1.15 x a single solve)
                    count = 1
            end = int(self.store.add(self.key, count))
            start = end - count
            return [int(i) for i in self.order[start:min(end, self.n)]] if start < self.n else []
        chunk = self.static[self.pos:self.pos + count]
        self.pos += len(chunk)
        return [int(i) for i in chunk]


class SweepEngine:
    """One mesh, one set of displacement conditions, one material on one GPU; load cases (target-force fields) are
    relaxed one after another through persistent contexts -- one or two at a time where the register/shared-memory
    resident CG kernels apply, in lockstep batches otherwise (grid barriers amortised)."""

    def __init__(self, coords, tets, bc_displacement, material, step, max_iter, conv_crit, device=0, batch=None, beams=None,
                 curve_bins=None, initial_displacements=None):
        from .solver import build_beams
        self.coords = np.ascontiguousarray(coords, dtype=np.float64)
        self.tets = np.ascontiguousarray(tets).astype(np.int32)
        self.movable = np.any(np.isnan(bc_displacement), axis=1)
        self.U0 = np.zeros_like(self.coords)
        self.U0[~self.movable] = np.asarray(bc_displacement)[~self.movable]
        if initial_displacements is not None:
            self.U0[self.movable] = np.asarray(initial_displacements, dtype=np.float64)[self.movable]
        self.beams = build_beams(300) if beams is None else beams
        self.mat = (material['K_0'], material['D_0'], material['L_S'], material['D_S'])
        self.step, self.max_iter, self.conv_crit = step, max_iter, conv_crit
        self.device, self.curve_bins = device, curve_bins
        self.contexts = {}
        self.totals = None
        ctx = self._context(1)
        if batch:
            self.batch = max(1, min(int(batch), 64))
        elif ctx.info()["cg_resident"]:
            # matrix on chip: one load case at a time -- or two side by side where the pair kernel applies (two CTAs per
            # SM, one system each, independent CG solves in one launch: info cg_fused == 4)
            self.batch = 1
            if not os.environ.get("JF_CG_NO_PAIR"):
                if self._context(2).info()["cg_fused"] == 4:
                    self.batch = 2
                else:
                    self.contexts.pop(2).close()
        else:
            self.batch = MAX_BATCH

    def _context(self, size):
        if size not in self.contexts:
            ctx = _lib.Context(self.coords, self.tets, self.movable, n_sys=size, beams=self.beams, device=self.device)
            ctx.set_material(*self.mat)
            if self.curve_bins is not None:
                ctx.set_curve_bins(self.coords, *self.curve_bins)
            self.contexts[size] = ctx
        return self.contexts[size]

    def relax_batch(self, forces_list, fetch_fields=True, keep_displacements=False):
        """Relax ``len(forces_list) <= batch`` load cases in lockstep.  Returns one dict per case: ``relrec, cg_iters,
        ms`` (device time of the batch divided by its size) and, as asked, ``U, forces, curve``."""
        size = len(forces_list)
        ctx = self._context(size)
        ctx.reset_stats()
        for s, f in enumerate(forces_list):
            ctx.set_targets(np.nan_to_num(f), s)
            if not keep_displacements:
                ctx.set_displacements(self.U0, s)
        relrec, cgit, _ = ctx.relax(self.step, self.max_iter, self.conv_crit)
        curves = ctx.curves() if self.curve_bins is not None else None
        st = ctx.stats()
        out = []
        for s in range(size):
            res = dict(relrec=relrec[s], cg_iters=cgit[s], ms=st["ms_relax"] / size)
            if fetch_fields:
                res.update(U=ctx.displacements(s), forces=ctx.forces(s))
            if curves is not None:
                res['curve'] = curves[s]
            out.append(res)
        st = ctx.stats()   # including the downloads
        self.totals = st if self.totals is None else {k: self.totals[k] + st[k] for k in st}
        return out

    def info(self):
        """Library info of the context the engine relaxes full batches with."""
        return self._context(self.batch).info()

    def close(self):
        for ctx in self.contexts.values():
            ctx.close()
        self.contexts = {}

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def relax_load_cases(coords, tets, bc_displacement, bc_forces_list, material, step, max_iter, conv_crit,
                     initial_displacements=None, device=0, batch=None, beams=None, curve_bins=None, fetch_fields=True,
                     warm_start=False):
    """Relax every load case in ``bc_forces_list`` (same mesh, same displacement conditions) and
    return a list of dicts ``U, forces, relrec, cg_iters`` in the same order, plus the summed
    library statistics.  ``curve_bins=(r_inner, x_edges)`` adds ``curve`` (radial median curve computed on the
    device); ``fetch_fields=False`` then skips the download of U and forces altogether.  ``warm_start=True``
    (opt-in, one load case at a time) starts every case from the relaxed displacements of the previous one, which
    stay on the device -- what the reference's ``get_initial`` intends (simulation.py:209-213, 227-233)."""
    n = len(bc_forces_list)
    results = []
    with SweepEngine(coords, tets, bc_displacement, material, step, max_iter, conv_crit, device=device,
                     batch=1 if warm_start else batch, beams=beams, curve_bins=curve_bins,
                     initial_displacements=initial_displacements) as eng:
        B = max(1, min(n, eng.batch))
        for a in range(0, n, B):
            results += eng.relax_batch(bc_forces_list[a:a + B], fetch_fields=fetch_fields,
                                       keep_displacements=warm_start and a > 0)
        totals = eng.totals
    return results, totals


def _write_case(folder, coords, tets, material, mat_par, value, bc, r_in, r_out, res, legacy_npz=True):
    from . import simulation as sim
    from .solver import save_solver_file
    os.makedirs(folder, exist_ok=True)
    sim.write_parameters(folder, material, value, bc, r_in, r_out, len(coords))
    np.savetxt(folder + "/relrec.txt", res['relrec'])
    movable = np.any(np.isnan(bc['displacement']), axis=1)
    save_solver_file(folder + "/solver.saenopy", coords, tets, res['U'], res['forces'], np.nan_to_num(bc['forces']), movable,
                     mat_par, res['relrec'])
    if legacy_npz:   # what the reference's last loader fallback reads (simulation.py:252-258, 289-291) whatever saenopy it runs on
        save_solver_file(folder + "/solver.npz", coords, tets, res['U'], res['forces'], np.nan_to_num(bc['forces']), movable,
                         mat_par, res['relrec'], layout="legacy")


def run_sweep(meshfile, outfolder, values, material, var_arg='pressure', max_iter=600, step=0.0033, conv_crit=0.01,
              batch=None, callback=None, r_inner=None, r_outer=None, x0=1, x1=50, n_bins=100, write=True,
              warm_start=False, schedule="dynamic", engine=None, legacy_npz=True):
    """Solve ``values`` (pressures), write the reference's folder layout and return the lookup table
    ``{'pressure','x','y'}`` (complete on every rank when ``torch.distributed`` is initialised, else this
    rank's rows).  Every folder also gets a legacy ``solver.npz`` (keys R, T, U, f; ``legacy_npz=False`` skips it): the
    reference's table builder falls back to it when the saenopy it runs on cannot read our ``solver.saenopy``
    (simulation.py:278-291).  ``schedule``: 'dynamic' (shared work counter, needs a process group; falls back to 'lpt'
    without one) or 'lpt' (static, by ``pressure_cost``).  ``warm_start`` implies a static ascending order on one
    rank's share.  ``engine``: a live ``SweepEngine`` for this mesh/material to reuse (its context stays resident)."""
    from . import simulation as sim
    from .solver import SemiAffineFiberMaterial

    rank, world = _env_rank_world()
    if world > 1 and _group_store() is None:
        warnings.warn("WORLD_SIZE=%d but torch.distributed is not initialised: this rank solves its static share and "
                      "returns a PARTIAL table (initialise the process group to gather it)" % world, RuntimeWarning)
    device = int(os.environ.get("LOCAL_RANK", "0")) if world > 1 else 0
    values = np.asarray(values, dtype=np.float64)
    coords, tets, r_in, r_out = sim.read_meshfile(meshfile, r_inner, r_outer)

    if write and rank == 0:
        os.makedirs(outfolder, exist_ok=True)
        np.savetxt(outfolder + '/' + var_arg + '-values.txt', values)

    x, x_center = sim.lookup_bins(x0, x1, n_bins)
    # the radius exactly as the table builder recovers it from parameters.txt (simulation.py:296-297)
    r_curve = (r_in * 1e6) * 10 ** -6
    costs = pressure_cost(values, material)
    own = engine is None
    if own:
        bc0 = sim.spherical_boundary_conditions(coords, r_in, r_out, float(values[0]) if len(values) else 1.0)
        engine = SweepEngine(coords, tets, bc0['displacement'], material, step, max_iter, conv_crit, device=device,
                             batch=1 if warm_start else batch, curve_bins=(r_curve, x))
    order = np.sort(np.arange(len(values))) if warm_start else work_order(values, costs, engine.batch)
    claimer = _Claimer(order, costs, rank, world, "lpt" if warm_start else schedule, batch=engine.batch)
    mat_par = SemiAffineFiberMaterial(material['K_0'], material['D_0'], material['L_S'], material['D_S']).serialize()
    tets64 = tets.astype(np.int64)
    done, curves, ms = [], [], []
    writer = ThreadPoolExecutor(max_workers=2) if write else None
    pending = []
    try:
        first = True
        while True:
            chunk = claimer.take(engine.batch)
            if not chunk:
                break
            bcs = [sim.spherical_boundary_conditions(coords, r_in, r_out, float(values[i])) for i in chunk]
            results = engine.relax_batch([b['forces'] for b in bcs], fetch_fields=write,
                                         keep_displacements=warm_start and not first)
            first = False
            for i, bc, res in zip(chunk, bcs, results):
                done.append(i)
                curves.append(res['curve'])   # radial median on the device (jf_curves.cu), bit-identical to the host formula
                ms.append(res['ms'])
                if write:
                    folder = outfolder + '/simulation' + str(int(i)).zfill(6)
                    pending.append(writer.submit(_write_case, folder, coords, tets64, material, mat_par, float(values[i]), bc,
                                                 r_in, r_out, res, legacy_npz))
                if callback is not None:
                    callback(int(i), len(values))
        for p in pending:
            p.result()   # re-raises what a writer thread raised
        stats = engine.totals
    finally:
        if writer is not None:
            writer.shutdown(wait=True)
        if own and engine is not None:
            engine.close()

    done = np.asarray(done, dtype=np.int64)
    srt = np.argsort(done, kind="stable")
    curves = np.asarray(curves, dtype=np.float64).reshape(len(done), n_bins)[srt]
    table = gather_table(done[srt], values[done[srt]], curves, len(values), n_bins, device=device,
                         extra=np.asarray(ms, dtype=np.float64)[srt])
    table['x'] = x_center
    table['stats'] = stats
    return table


def gather_table(indices, pressures, curves, n_total, n_bins, device=None, extra=None):
    """Collect every rank's curve rows into the full table.  With an initialised process group
    this is one ``all_gather`` (NCCL over NVLink on GPUs, gloo in the CPU tests) of a fixed-size
    padded block per rank; without one it returns the local rows.  ``extra``: one more number per row (device
    milliseconds of the solve) that travels along as ``table['ms']``."""
    indices = np.asarray(indices, dtype=np.int64)
    k = len(indices)
    extra = np.zeros(k) if extra is None else np.asarray(extra, dtype=np.float64)
    try:
        import torch
        import torch.distributed as dist
        have = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    except Exception:
        have = False
    if not have:
        return {'pressure': np.asarray(pressures, dtype=np.float64), 'y': np.asarray(curves), 'index': indices, 'ms': extra,
                'rank': np.zeros(k, dtype=np.int64)}
    world = dist.get_world_size()
    cap = n_total   # a dynamic schedule may hand one rank (almost) everything
    block = torch.full((cap, n_bins + 4), float('nan'), dtype=torch.float64)
    if k:
        block[:k, 0] = torch.from_numpy(indices.astype(np.float64))
        block[:k, 1] = torch.from_numpy(np.asarray(pressures, dtype=np.float64))
        block[:k, 2] = torch.from_numpy(extra)
        block[:k, 3] = float(dist.get_rank())
        block[:k, 4:] = torch.from_numpy(np.asarray(curves, dtype=np.float64).reshape(k, n_bins))
    block[k:, 0] = -1.0
    if dist.get_backend() == "nccl":
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", torch.cuda.current_device()))
        dev = torch.device("cuda", int(device))
    else:
        dev = torch.device("cpu")
    block = block.to(dev)
    parts = [torch.empty_like(block) for _ in range(world)]
    dist.all_gather(parts, block)
    allrows = torch.cat(parts).cpu().numpy()
    allrows = allrows[allrows[:, 0] >= 0]
    allrows = allrows[np.argsort(allrows[:, 0], kind="stable")]
    return {'pressure': allrows[:, 1].copy(), 'y': allrows[:, 4:].copy(), 'index': allrows[:, 0].astype(np.int64),
            'ms': allrows[:, 2].copy(), 'rank': allrows[:, 3].astype(np.int64)}
