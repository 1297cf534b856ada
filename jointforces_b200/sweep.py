"""Pressure sweep: many load cases on one mesh, sharded over GPUs, relaxed in lockstep batches.

Replaces the process-per-pressure loop of ``/root/reference/jointforces/simulation.py:204-246``.
The solves are independent (every one is a cold start), so ranks share
nothing on the data path; the only collective is the final gather of the (n, 100) displacement
curves over NCCL.  One process per GPU: run under ``torchrun``/``torch.distributed.run`` for
several GPUs (RANK / WORLD_SIZE / LOCAL_RANK are read from the environment), or plainly for one.
"""
from __future__ import annotations

import os

import numpy as np

from . import _lib

MAX_BATCH = 16  # systems relaxed in lockstep per context (library limit is 64)


def _env_rank_world():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_indices(values, rank, world):
    """Largest value first, dealt round-robin: cost grows steeply with pressure (more Newton and
    CG iterations), so this balances better than contiguous blocks.  Returned ascending."""
    order = np.argsort(-np.asarray(values, dtype=np.float64), kind="stable")
    return np.sort(order[rank::world])


def relax_load_cases(coords, tets, bc_displacement, bc_forces_list, material, step, max_iter, conv_crit,
                     initial_displacements=None, device=0, batch=None, beams=None, curve_bins=None, fetch_fields=True,
                     warm_start=False):
    """Relax every load case in ``bc_forces_list`` (same mesh, same displacement conditions) and
    return a list of dicts ``U, forces, relrec, cg_iters`` in the same order, plus the summed
    library statistics.  ``curve_bins=(r_inner, x_edges)`` adds ``curve`` (radial median curve computed on the
    device); ``fetch_fields=False`` then skips the download of U and forces altogether.  ``warm_start=True``
    (opt-in, one load case at a time) starts every case from the relaxed displacements of the previous one, which
    stay on the device -- what the reference's ``get_initial`` intends (simulation.py:209-213, 227-233)."""
    from .solver import build_beams
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    tets_i = np.ascontiguousarray(tets).astype(np.int32)
    movable = np.any(np.isnan(bc_displacement), axis=1)
    U0 = np.zeros_like(coords)
    U0[~movable] = bc_displacement[~movable]
    if initial_displacements is not None:
        U0[movable] = np.asarray(initial_displacements, dtype=np.float64)[movable]
    if beams is None:
        beams = build_beams(300)
    n = len(bc_forces_list)
    mat = (material['K_0'], material['D_0'], material['L_S'], material['D_S'])
    results = [None] * n
    totals = None
    contexts = {}
    if warm_start:
        batch = 1
    if batch:
        B = max(1, min(n, int(batch), 64))
    else:
        # meshes small enough for the resident CG kernel (matrix kept on chip for a whole CG solve) run fastest
        # one load case at a time; larger ones amortise the grid barriers over a lockstep batch
        contexts[1] = _lib.Context(coords, tets_i, movable, n_sys=1, beams=beams, device=device)
        contexts[1].set_material(*mat)
        if curve_bins is not None:
            contexts[1].set_curve_bins(coords, *curve_bins)
        B = 1 if contexts[1].info()["cg_resident"] else max(1, min(n, MAX_BATCH))
    try:
        for a in range(0, n, B):
            idx = list(range(a, min(n, a + B)))
            size = len(idx)
            if size not in contexts:
                contexts[size] = _lib.Context(coords, tets_i, movable, n_sys=size, beams=beams, device=device)
                contexts[size].set_material(*mat)
                if curve_bins is not None:
                    contexts[size].set_curve_bins(coords, *curve_bins)
            ctx = contexts[size]
            ctx.reset_stats()
            for s, i in enumerate(idx):
                ctx.set_targets(np.nan_to_num(bc_forces_list[i]), s)
                if not (warm_start and a > 0):
                    ctx.set_displacements(U0, s)
            relrec, cgit, _ = ctx.relax(step, max_iter, conv_crit)
            curves = ctx.curves() if curve_bins is not None else None
            for s, i in enumerate(idx):
                results[i] = dict(relrec=relrec[s], cg_iters=cgit[s])
                if fetch_fields:
                    results[i].update(U=ctx.displacements(s), forces=ctx.forces(s))
                if curves is not None:
                    results[i]['curve'] = curves[s]
            st = ctx.stats()
            totals = st if totals is None else {k: totals[k] + st[k] for k in st}
    finally:
        for ctx in contexts.values():
            ctx.close()
    return results, totals


def run_sweep(meshfile, outfolder, values, material, var_arg='pressure', max_iter=600, step=0.0033, conv_crit=0.01,
              batch=None, callback=None, r_inner=None, r_outer=None, x0=1, x1=50, n_bins=100, write=True,
              warm_start=False):
    """Solve ``values`` (pressures) on this rank's share, write the reference's folder layout and
    return the lookup table ``{'pressure','x','y'}`` (complete on every rank when
    ``torch.distributed`` is initialised, else this rank's rows)."""
    from . import simulation as sim
    from .solver import save_solver_file, SemiAffineFiberMaterial

    rank, world = _env_rank_world()
    device = int(os.environ.get("LOCAL_RANK", "0")) if world > 1 else 0
    values = np.asarray(values, dtype=np.float64)
    coords, tets, r_in, r_out = sim.read_meshfile(meshfile, r_inner, r_outer)

    if write and rank == 0:
        os.makedirs(outfolder, exist_ok=True)
        np.savetxt(outfolder + '/' + var_arg + '-values.txt', values)

    mine = shard_indices(values, rank, world)
    bcs = [sim.spherical_boundary_conditions(coords, r_in, r_out, float(values[i])) for i in mine]
    results, stats = ([], None)
    x, x_center = sim.lookup_bins(x0, x1, n_bins)
    # the radius exactly as the table builder recovers it from parameters.txt (simulation.py:296-297)
    r_curve = (r_in * 1e6) * 10 ** -6
    if len(mine):
        results, stats = relax_load_cases(coords, tets, bcs[0]['displacement'], [b['forces'] for b in bcs], material, step,
                                          max_iter, conv_crit, device=device, batch=batch, curve_bins=(r_curve, x),
                                          fetch_fields=write, warm_start=warm_start)
    curves = np.full((len(mine), n_bins), np.nan)
    mat_par = SemiAffineFiberMaterial(material['K_0'], material['D_0'], material['L_S'], material['D_S']).serialize()
    for j, (i, bc, res) in enumerate(zip(mine, bcs, results)):
        curves[j] = res['curve']   # radial median on the device (jf_curves.cu), bit-identical to the host formula
        if write:
            folder = outfolder + '/simulation' + str(int(i)).zfill(6)
            os.makedirs(folder, exist_ok=True)
            sim.write_parameters(folder, material, float(values[i]), bc, r_in, r_out, len(coords))
            np.savetxt(folder + "/relrec.txt", res['relrec'])
            movable = np.any(np.isnan(bc['displacement']), axis=1)
            save_solver_file(folder + "/solver.saenopy", coords, tets.astype(np.int64), res['U'], res['forces'],
                             np.nan_to_num(bc['forces']), movable, mat_par, res['relrec'])
        if callback is not None:
            callback(int(i), len(values))

    table = gather_table(mine, values[mine], curves, len(values), n_bins)
    table['x'] = x_center
    table['stats'] = stats
    return table


def gather_table(indices, pressures, curves, n_total, n_bins):
    """Collect every rank's curve rows into the full table.  With an initialised process group
    this is one ``all_gather`` (NCCL over NVLink on GPUs, gloo in the CPU tests) of a fixed-size
    padded block per rank; without one it returns the local rows."""
    indices = np.asarray(indices, dtype=np.int64)
    try:
        import torch
        import torch.distributed as dist
        have = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    except Exception:
        have = False
    if not have:
        return {'pressure': np.asarray(pressures, dtype=np.float64), 'y': np.asarray(curves), 'index': indices}
    world = dist.get_world_size()
    cap = (n_total + world - 1) // world
    block = torch.full((cap, n_bins + 2), float('nan'), dtype=torch.float64)
    k = len(indices)
    if k:
        block[:k, 0] = torch.from_numpy(indices.astype(np.float64))
        block[:k, 1] = torch.from_numpy(np.asarray(pressures, dtype=np.float64))
        block[:k, 2:] = torch.from_numpy(np.asarray(curves, dtype=np.float64))
    block[k:, 0] = -1.0
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    block = block.to(dev)
    parts = [torch.empty_like(block) for _ in range(world)]
    dist.all_gather(parts, block)
    allrows = torch.cat(parts).cpu().numpy()
    allrows = allrows[allrows[:, 0] >= 0]
    allrows = allrows[np.argsort(allrows[:, 0], kind="stable")]
    return {'pressure': allrows[:, 1].copy(), 'y': allrows[:, 2:].copy(), 'index': allrows[:, 0].astype(np.int64)}
