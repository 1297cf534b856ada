#!/usr/bin/env python
"""Benchmark of the relaxation hot path (contract: see the task statement / DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg3|cfg2|cfg1|cfg4|cfg5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Default workload ``cfg3`` = BASELINE configs[2], the lookup-table sweep the metric "sims/hour at 1/2/4/8
B200" is quoted on: collagen12 on the configs[1] mesh (17,612 nodes / 102,744 tets), pressures from
``np.logspace(-1, 4, 150)``.  One "step" relaxes every fifth pressure of that table (30 solves, all three
material regimes; five consecutive steps cover the table once) and gathers the 30 curve rows.  STRONG scaling:
the K steps of the timed region are ONE work list (K x 30 load cases, step after step) that the N ranks draw from
a shared counter in the process group's store -- a rank that is done with step j goes on with step j+1 -- and the rows
are gathered over NCCL at the end, as for a real 150-pressure table (one gather per table).

  value   sims/hour, whole job, device time: per step the max over ranks of the CUDA-event time of its
          relaxations (jf_relax), mesh structure resident in HBM
  e2e     the same metric through the public call ``run_sweep`` (what ``jf.simulation.distribute_solver``
          runs): mesh file on disk -> structure build -> H2D -> relaxations -> D2H -> the reference's
          output folders written -> NCCL gather; wall clock, max over ranks

``--workload cfg2`` (cfg1/cfg4/cfg5) times single solves instead: one step = one complete relaxation on each
rank's GPU (weak scaling); at N=1 the default run adds that number for configs[1] as ``single_solve``.

``--impl reference``: the reference's CPU implementation of the same workload -- the oracle's C port of the
saenopy algorithm (saenopy itself cannot be installed here; if ``import saenopy`` works on the
box it is used instead), run the way the reference runs a sweep (simulation.py:186-187, 218-246): several
solver processes alive at once, every one a COMPLETE solve, nothing extrapolated.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (BASELINE config text, length_factor, lf_iso override, material name, pressure)
    "cfg1": ("configs[0]: spherical_inclusion lf=0.2 (lf_iso 0.2), linear K_0=2364, 100 Pa", None, 0.2, "linear2364", 100.0),
    "cfg2": ("configs[1]: spherical_inclusion lf=0.05 (lf_iso 0.1667), collagen12, 100 Pa, single solve", 0.05, None, "collagen12", 100.0),
    "cfg3": ("configs[2]: collagen12 lookup-table sweep on the configs[1] mesh, np.logspace(-1,4,150); one step = every 5th "
             "pressure (30 solves), offset rotating with the step; the steps of the timed region form one work list", 0.05, None, "collagen12", None),
    "cfg4": ("configs[3]: spherical_inclusion lf=0.03 (lf_iso 0.1), collagen24, 10 kPa", 0.03, None, "collagen24", 10000.0),
    "cfg5": ("configs[4]: spherical_inclusion lf=0.015 (lf_iso 0.05), collagen06, 100 Pa", 0.015, None, "collagen06", 100.0),
}
SOLVER = dict(step=0.0033, max_iter=600, conv_crit=0.01)   # reference defaults, simulation.py:81
TABLE = np.logspace(-1, 4, 150)                             # README.md:101-108
STRIDE = 5                                                  # pressures per step = 150 / STRIDE
R_IN, R_OUT = 100 * 10 ** -6, 10000 * 10 ** -6


def step_indices(j):
    return np.arange(j % STRIDE, len(TABLE), STRIDE)


def material_of(name):
    from jointforces_b200 import materials
    m = materials.linear(2364) if name == "linear2364" else getattr(materials, name)
    return m, (float(m['K_0']), float(m['D_0']), float(m['L_S']), float(m['D_S']))


def build_case(workload):
    from jointforces_b200.mesh import generate_spherical_inclusion
    from jointforces_b200.simulation import spherical_boundary_conditions
    from jointforces_b200.solver import build_beams
    text, lf, lf_iso, matname, pressure = WORKLOADS[workload]
    coords, tets = generate_spherical_inclusion(length_factor=lf or 0.05, lf_iso=lf_iso)
    bc = spherical_boundary_conditions(coords, R_IN, R_OUT, pressure if pressure is not None else 1.0)
    movable = np.any(np.isnan(bc["displacement"]), axis=1)
    mat_dict, mat = material_of(matname)
    return dict(text=text, coords=coords, tets=tets.astype(np.int32), movable=movable, ft=np.nan_to_num(bc["forces"]),
                bc=bc, mat=mat, mat_dict=mat_dict, matname=matname, pressure=pressure, r_in=R_IN, beams=build_beams(300),
                length_factor=lf or 0.05, lf_iso=lf_iso)


def shared_config(case, workload):
    """The part of the JSON line both arms print identically."""
    cfg = {"workload": case["text"], "n_nodes": int(len(case["coords"])), "n_tets": int(len(case["tets"])),
           "material": case["matname"], "solver": SOLVER}
    if workload == "cfg3":
        cfg.update(pressures_per_step=int(len(step_indices(0))), table="np.logspace(-1, 4, 150)[(step % 5)::5]")
    else:
        cfg.update(pressure_pa=case["pressure"])
    return cfg


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        time.sleep(0.05)
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 9 for i in range(4) if r[5 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


def cg_bytes_per_iteration(info):
    """Algorithmic bytes of one CG iteration of one system, SURVEY.md §8(d): every non-zero 3x3
    block once (72 B) with its column index (4 B); per free block row a row pointer (4 B), the x
    read and y write of the SpMV (24 + 24 B) and nine 24-byte vector passes of the fused updates."""
    return info["n_blocks"] * 76 + info["n_free"] * (4 + 24 + 24 + 9 * 24)


def cg_kernel_name(info):
    if info.get("cg_fused") == 4:
        return "jf_cg_pair_kernel"
    if info.get("cg_fused") == 1:
        return "jf_cg_fused_kernel"
    if info.get("cg_fused") == 2:
        return "jf_cg_fused_stream_kernel"
    if info["cg_resident"] == 2:
        return "jf_cg_regres_kernel"
    if info["cg_resident"] == 1:
        return "jf_cg_resident_kernel"
    return "jf_cg_streamhalo_kernel" if info.get("cg_stream_halo") else "jf_cg_stream_kernel"


def roofline_of(info, st, peaks, peak_kind, workload):
    """The dominant kernel (CG).  Where the matrix is register/shared-memory resident the kernel is bound by the latency
    of its per-iteration all-reduce, not by HBM: `achieved` is then a RATE of algorithmic bytes (what an HBM-streaming
    kernel would have to sustain for the same iteration time), `traffic` the DRAM bytes ncu measured for one launch."""
    by_iter = cg_bytes_per_iteration(info)
    achieved = by_iter * st["cg_iterations"] / (st["ms_cg"] * 1e-3) / 1e9
    traffic = None
    tname = "cfg3_pair" if info.get("cg_fused") == 4 else ("cfg2" if workload == "cfg3" else workload)
    tpath = os.path.join(ROOT, "profiles", "cg_traffic_%s.json" % tname)
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    resident = bool(info["cg_resident"]) or info.get("cg_fused") == 4
    matrix_mb = info["n_blocks"] * 72 / 1e6
    return {"bound": "latency" if resident else "hbm", "nominal_bound": "hbm", "kernel": cg_kernel_name(info), "achieved": achieved,
            "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": traffic,
            "peak_kind": "of " + peak_kind, "algorithmic_bytes_per_launch": by_iter * st["cg_iterations"] / max(1, st["cg_launches"]),
            "us_per_cg_iteration": st["ms_cg"] * 1e3 / max(1, st["cg_iterations"]),
            "achieved_is": ("a rate of algorithmic bytes: the %.0f MB matrix of a system is staged once per CG solve into registers and "
                            "tensor / shared memory, DRAM is idle (`traffic`); the iteration is bound by its grid-wide all-reduce%s. "
                            "roofline_large is the HBM-bound regime"
                            % (matrix_mb, " (two systems per launch, two CTAs per SM: us_per_cg_iteration is per system-iteration)"
                               if info.get("cg_fused") == 4 else "")) if resident else "HBM traffic (matrix streamed)"}


def hbm_regime_leg(device, peaks):
    """The same CG kernel family where the matrix cannot be cache resident (BASELINE configs[4] size):
    a structured graded shell mesh (502,968 nodes / 2,989,980 tets), one evaluation, CG launches of exactly
    50 iterations timed with the library's CUDA events."""
    from jointforces_b200 import _lib, materials
    from jointforces_b200.mesh import generate_spherical_inclusion
    from jointforces_b200.simulation import spherical_boundary_conditions
    from jointforces_b200.solver import build_beams
    coords, tets = generate_spherical_inclusion(lf_iso=0.05, method="icoshell")
    bc = spherical_boundary_conditions(coords, R_IN, R_OUT, 100.0)
    movable = np.any(np.isnan(bc["displacement"]), axis=1)
    m = materials.collagen06
    with _lib.Context(coords, tets.astype(np.int32), movable, n_sys=1, beams=build_beams(300), device=device) as ctx:
        ctx.set_material(m['K_0'], m['D_0'], m['L_S'], m['D_S'])
        ctx.set_targets(np.nan_to_num(bc["forces"]))
        ctx.set_displacements(np.zeros_like(coords))
        ctx.evaluate()
        for _ in range(3):
            ctx.cg(1e-30, 50)
        ctx.reset_stats()
        reps = 5
        for _ in range(reps):
            ctx.cg(1e-30, 50)
        st, info = ctx.stats(), ctx.info()
    by = cg_bytes_per_iteration(info)
    achieved = by * 50 * reps / (st["ms_cg"] * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "cg_traffic_large.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    return {"bound": "hbm", "kernel": cg_kernel_name(info), "workload": "configs[4]-size icoshell mesh, %d nodes / %d tets, "
            "collagen06; matrix %.0f MB" % (info["n_nodes"], info["n_tets"], info["n_blocks"] * 72 / 1e6),
            "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
            "us_per_cg_iteration": st["ms_cg"] * 1e3 / (50 * reps), "algorithmic_bytes_per_launch": by * 50,
            "traffic": traffic}


# ---------------------------------------------------------------------------------------------------------------------
# CPU side: complete solves with the oracle (or saenopy where it can be imported)
# ---------------------------------------------------------------------------------------------------------------------
def _cpu_solve_worker(args):
    """One complete relaxation in a worker process.  Returns (seconds, newton steps, cg iterations, kind)."""
    npz, pressure, mat, threads, want_saenopy = args
    sys.path.insert(0, ROOT)
    d = np.load(npz)
    coords, tets = d["coords"], d["tets"]
    from jointforces_b200.simulation import spherical_boundary_conditions
    bc = spherical_boundary_conditions(coords, R_IN, R_OUT, pressure)
    t0 = time.time()
    if want_saenopy:
        import saenopy                                             # the genuine reference path (simulation.py:100-167)
        from saenopy.materials import SemiAffineFiberMaterial
        M = saenopy.Solver()
        M.set_material_model(SemiAffineFiberMaterial(*mat))
        M.set_nodes(coords)
        M.set_tetrahedra(tets.astype(np.float64))
        M.set_boundary_condition(bc["displacement"], bc["forces"])
        rr = M.solve_boundarycondition(step_size=SOLVER["step"], max_iterations=SOLVER["max_iter"], rel_conv_crit=SOLVER["conv_crit"])
        return time.time() - t0, len(rr) - 1, -1, "reference"
    from oracle import c_oracle
    from oracle.saeno_oracle import build_beams
    movable = np.any(np.isnan(bc["displacement"]), axis=1)
    r = c_oracle.relax(coords, tets.astype(np.int32), movable, np.nan_to_num(bc["forces"]), np.zeros_like(coords),
                       build_beams(300), mat, SOLVER["step"], SOLVER["max_iter"], SOLVER["conv_crit"], nthreads=threads)
    return time.time() - t0, int(r["n_iter"]), int(r["cg_iters"].sum()), "port"


def saenopy_importable():
    try:
        import saenopy  # noqa: F401
        return True
    except Exception:
        return False


class CpuPool:
    """The reference's sweep shape (simulation.py:186-187, 218-246): `procs` solver processes alive at once, one
    pressure each; here every process is the oracle's OpenMP C port on cores/procs threads."""

    def __init__(self, case, procs=None):
        self.cores = os.cpu_count() or 1
        self.procs = procs or max(1, min(8, self.cores // 8))
        self.threads = max(1, self.cores // self.procs)
        self.saenopy = saenopy_importable()
        self.case = case
        self.tmp = tempfile.mkdtemp(prefix="jf_cpu_")
        self.npz = os.path.join(self.tmp, "mesh.npz")
        np.savez(self.npz, coords=case["coords"], tets=case["tets"])
        import multiprocessing as mp
        self.pool = mp.get_context("spawn").Pool(self.procs)

    def wave(self, pressures):
        """Solve `pressures` (<= procs of them) concurrently; returns (wall seconds, per-solve tuples)."""
        t0 = time.time()
        out = self.pool.map(_cpu_solve_worker, [(self.npz, float(p), self.case["mat"], self.threads, self.saenopy) for p in pressures])
        return time.time() - t0, out

    def close(self):
        self.pool.close()
        self.pool.join()
        shutil.rmtree(self.tmp, ignore_errors=True)

    def describe(self):
        return ("saenopy %s" % __import__("saenopy").__version__) if self.saenopy else "oracle/saeno_oracle.c (C port, OpenMP)"


def wave_pressures(case, workload, j, procs):
    """The pressures one CPU wave solves: spread evenly over the step's subset (cheap and expensive cases alike)."""
    if workload != "cfg3":
        return [case["pressure"]] * procs
    idx = step_indices(j)
    pick = idx[np.round(np.linspace(0, len(idx) - 1, procs + 2)[1:-1]).astype(int)] if procs < len(idx) else idx
    return [float(TABLE[i]) for i in pick]


def run_reference(args, rank, world):
    """--impl reference (rank 0 only): complete CPU solves, several processes alive at once, nothing extrapolated.
    The number of timed steps is capped by a time budget; `steps` reports what was actually timed."""
    if rank != 0:
        return
    case = build_case(args.workload)
    pool = CpuPool(case, args.cpu_procs)
    budget = args.cpu_budget
    try:
        t_start = time.time()
        walls, solves = [], []
        j = 0
        while j < args.steps and (j == 0 or time.time() - t_start + (np.mean(walls) if walls else 0) < budget):
            wall, out = pool.wave(wave_pressures(case, args.workload, j, pool.procs))
            walls.append(wall)
            solves += out
            j += 1
    finally:
        pool.close()
    n_solves = len(solves)
    value = n_solves * 3600.0 / float(np.sum(walls))
    kind = solves[0][3]
    sample = ("%d complete solves (%d waves of %d processes x %d threads, %s; %.1f s per wave; Newton steps %s)"
              % (n_solves, len(walls), pool.procs, pool.threads, pool.describe(), float(np.mean(walls)),
                 sorted({s[1] for s in solves})))
    line = {
        "impl": "reference", "metric": "lookup-table sims/hour", "value": value, "unit": "sims/hour", "n_gpus": world,
        "steps": len(walls), "steps_requested": args.steps, "warmup": 0, "ms_per_step": 1e3 * float(np.mean(walls)),
        "higher_is_better": True, "scaling": "strong" if args.workload == "cfg3" else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": shared_config(case, args.workload),
        "note": "every timed step is one wave of complete CPU solves; the number of waves is capped by a %d s budget "
                "(--cpu-budget), no warm-up wave (nothing to warm: every solve is a cold start)" % budget,
        "cpu_baseline": {"value": value, "unit": "sims/hour", "cores": pool.cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "sims/hour", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------
# GPU side
# ---------------------------------------------------------------------------------------------------------------------
def single_solve_legs(case, local, K, W, flush, barrier, world, rank, dist, torch):
    """Device-timed and end-to-end single relaxations of `case` (one per rank).  Returns a dict of numbers."""
    from jointforces_b200 import _lib
    from jointforces_b200.simulation import radial_median_curve, lookup_bins
    from jointforces_b200.sweep import gather_table
    x_edges, _ = lookup_bins()
    ctx = _lib.Context(case["coords"], case["tets"], case["movable"], n_sys=1, beams=case["beams"], device=local)
    ctx.set_material(*case["mat"])
    ctx.set_targets(case["ft"])
    ctx.set_displacements(np.zeros_like(case["coords"]))
    info = ctx.info()

    def device_step():
        flush.fill_(1)
        torch.cuda.synchronize()
        ctx.restore_displacements()
        return ctx.relax(SOLVER["step"], SOLVER["max_iter"], SOLVER["conv_crit"])

    for _ in range(W):
        device_step()
    ctx.reset_stats()
    barrier()
    t0 = time.time()
    for _ in range(K):
        relrec, cgit, n_iter = device_step()
    barrier()
    wall = time.time() - t0
    st = ctx.stats()
    ctx.close()
    ms_dev = torch.tensor([st["ms_relax"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms_dev, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms_dev.item()) / K

    def e2e_step():
        out = _lib.solve_boundarycondition(case["coords"], case["tets"], case["movable"], case["ft"],
                                           np.zeros_like(case["coords"]), case["mat"], SOLVER["step"], SOLVER["max_iter"],
                                           SOLVER["conv_crit"], beams=case["beams"], device=local)
        curve = radial_median_curve(case["coords"], out["U"], case["r_in"], x_edges)
        gather_table([rank], [case["pressure"]], curve[None, :], world, len(curve), device=local)
        return out

    e2e_step()
    Ke = min(K, 5)
    barrier()
    t0 = time.time()
    for _ in range(Ke):
        flush.fill_(1)
        out = e2e_step()
    barrier()
    e2e_ms = torch.tensor([(time.time() - t0) * 1e3 / Ke], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    return dict(info=info, st=st, ms_per_step=ms_per_step, wall_ms=wall * 1e3 / K, e2e_ms=float(e2e_ms.item()), e2e_steps=Ke,
                h2d=int(out["stats"]["bytes_h2d"]), d2h=int(out["stats"]["bytes_d2h"]), newton=int(n_iter[0]), cg=int(cgit[0].sum()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-procs", type=int, default=0, help="CPU solver processes alive at once (0: cores // 8, at most 8)")
    ap.add_argument("--cpu-budget", type=float, default=200.0, help="seconds of CPU solves in the reference arm")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-hbm-leg", action="store_true", help="skip the large-mesh (HBM-bound) CG roofline leg")
    ap.add_argument("--no-single", action="store_true", help="skip the configs[1] single-solve leg of the default run")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    # the library mirrors the reference's prints (mesh geometry notices ...): keep stdout for the one JSON line
    real_stdout = sys.stdout
    sys.stdout = sys.stderr

    import torch
    import torch.distributed as dist
    from jointforces_b200 import sweep as jsweep
    from jointforces_b200.simulation import lookup_bins, spherical_boundary_conditions
    from jointforces_b200.mesh import write_msh22

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    case = build_case(args.workload)
    W = max(args.warmup, 3)
    K = args.steps
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    peaks, peak_kind = measured_peaks()
    line = {"metric": "lookup-table sims/hour", "unit": "sims/hour", "n_gpus": world, "steps": K, "warmup": W,
            "higher_is_better": True, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": shared_config(case, args.workload)}

    if args.workload != "cfg3":
        sampler = ClockSampler(local)
        sampler.start()
        r = single_solve_legs(case, local, K, W, flush, barrier, world, rank, dist, torch)
        clocks = sampler.stop()
        if rank != 0:
            if world > 1:
                dist.destroy_process_group()
            return
        st, info = r["st"], r["info"]
        line.update(value=world * 3600.0 / (r["ms_per_step"] * 1e-3), ms_per_step=r["ms_per_step"], scaling="weak", clocks=clocks,
                    e2e={"value": world * 3600.0 / (r["e2e_ms"] * 1e-3), "unit": "sims/hour", "h2d_bytes_per_step": r["h2d"],
                         "d2h_bytes_per_step": r["d2h"], "ms_per_step": r["e2e_ms"], "steps": r["e2e_steps"]},
                    gpu_launches=int(st["kernel_launches"]),
                    submetrics={"tet_updates_per_s": st["tet_updates"] / (st["ms_element"] * 1e-3),
                                "cg_iters_per_s": st["cg_iterations"] / (st["ms_cg"] * 1e-3), "ms_element": st["ms_element"] / K,
                                "ms_assembly": st["ms_assembly"] / K, "ms_cg": st["ms_cg"] / K, "wall_ms_per_step": r["wall_ms"],
                                "newton_steps": r["newton"], "cg_iterations_per_solve": r["cg"]},
                    roofline=roofline_of(info, st, peaks, peak_kind, args.workload),
                    notes={"l2": "flushed between steps (256 MB write)", "parallelism": "independent solves, 1 per GPU",
                           "nonzero_blocks": info["n_blocks"], "directions": "%d folded to %d" % (info["n_beams"], info["n_dirs"])})
    else:
        # ---- the sweep: device-timed leg through a resident engine ---------------------------------------------------
        mat_dict = case["mat_dict"]
        x_edges, _ = lookup_bins()
        r_curve = (R_IN * 1e6) * 10 ** -6
        engine = jsweep.SweepEngine(case["coords"], case["tets"], case["bc"]["displacement"], mat_dict, SOLVER["step"],
                                    SOLVER["max_iter"], SOLVER["conv_crit"], device=local, curve_bins=(r_curve, x_edges))
        info = engine.info()
        costs_all = jsweep.pressure_cost(TABLE, mat_dict)

        def device_region(j0, nsteps):
            """`nsteps` steps as ONE work list (step after step, the most expensive case of a step first), drawn from the
            shared counter: a rank that finishes its last case of step j goes on with step j+1, so the only idle time is
            at the end of the region -- the way a real table (150 cases, one gather) is scheduled.  A rank writes the
            L2-flush buffer whenever it enters a new step.  Returns (max, min) over ranks of the device time and the
            number of cases this rank relaxed."""
            flat, step_of = [], []
            for j in range(j0, j0 + nsteps):
                idx = step_indices(j)
                flat += [int(i) for i in idx[jsweep.work_order(TABLE[idx], costs_all[idx], engine.batch)]]
                step_of += [j] * len(idx)
            flat = np.asarray(flat)
            claimer = jsweep._Claimer(np.arange(len(flat)), costs_all[flat], rank, world, "dynamic", batch=engine.batch)
            ms, rows, done, last_step = 0.0, [], [], None
            while True:
                chunk = claimer.take(engine.batch)
                if not chunk:
                    break
                if step_of[chunk[0]] != last_step:
                    last_step = step_of[chunk[0]]
                    flush.fill_(1)
                    torch.cuda.synchronize()
                cases = [int(flat[q]) for q in chunk]
                forces = [spherical_boundary_conditions(case["coords"], R_IN, R_OUT, float(TABLE[i]))["forces"] for i in cases]
                for q, res in zip(chunk, engine.relax_batch(forces, fetch_fields=False)):
                    ms += res["ms"]
                    rows.append(res["curve"])
                    done.append(q)
            jsweep.gather_table(done, TABLE[flat[done]] if done else [], np.asarray(rows).reshape(len(done), -1), len(flat),
                                len(x_edges) - 1, device=local)
            t = torch.tensor([ms, -ms], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)      # (max, -min) of the ranks' device time
            return float(t[0].item()), -float(t[1].item()), len(done)

        device_region(0, W)
        engine.totals = None
        sampler = ClockSampler(local)
        barrier()
        sampler.start()
        t_wall0 = time.time()
        ms_max, ms_min, mine = device_region(W, K)
        barrier()
        wall = time.time() - t_wall0
        clocks = sampler.stop()
        st = engine.totals or {k: 0.0 for k in ("kernel_launches", "tet_updates", "cg_iterations", "cg_launches")}
        st.update({k: st.get(k) or 1e-9 for k in ("ms_element", "ms_assembly", "ms_cg")})   # (a rank that drew no case)
        engine.close()
        S = len(step_indices(0))
        value = K * S * 3600.0 / (ms_max * 1e-3)
        launches = torch.tensor([float(st.get("kernel_launches", 0))], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(launches, op=dist.ReduceOp.SUM)

        # ---- end to end through run_sweep: mesh file in, folders out ----------------------------------------------
        tmp = tempfile.mkdtemp(prefix="jf_bench_") if rank == 0 else None
        if world > 1:
            box = [tmp]
            dist.broadcast_object_list(box, src=0)
            tmp = box[0]
        meshfile = os.path.join(tmp, "mesh.msh")
        if rank == 0:
            write_msh22(meshfile, case["coords"], case["tets"], 100, 10000, case["length_factor"])
        barrier()

        def e2e_call(j0, nsteps):
            """ONE run_sweep call (the call a user makes) over the pressures of `nsteps` steps -- the whole 150-pressure
            table when nsteps = 5."""
            out = os.path.join(tmp, "sweep%d" % j0)
            values = np.concatenate([TABLE[step_indices(j)] for j in range(j0, j0 + nsteps)])
            flush.fill_(1)
            barrier()
            t0 = time.time()
            table = jsweep.run_sweep(meshfile, out, values, mat_dict, max_iter=SOLVER["max_iter"], step=SOLVER["step"],
                                     conv_crit=SOLVER["conv_crit"])
            barrier()
            dt = time.time() - t0
            if rank == 0:
                shutil.rmtree(out, ignore_errors=True)
            return dt, table

        e2e_call(0, 1)
        Ke = min(K, STRIDE)
        e2e_s, table = e2e_call(W, Ke)
        h2d = table["stats"]["bytes_h2d"] if table["stats"] else 0
        d2h = table["stats"]["bytes_d2h"] if table["stats"] else 0
        tt = torch.tensor([e2e_s, h2d, d2h], dtype=torch.float64, device="cuda")
        if world > 1:
            mx = tt[:1].clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(tt, op=dist.ReduceOp.SUM)
            tt[0] = mx[0]
        e2e_s, h2d, d2h = (float(v) for v in tt.tolist())
        if rank == 0:
            shutil.rmtree(tmp, ignore_errors=True)
        if rank != 0:
            dist.destroy_process_group()
            return
        line.update(value=value, ms_per_step=ms_max / K, scaling="strong", clocks=clocks,
                    e2e={"value": Ke * S * 3600.0 / e2e_s, "unit": "sims/hour", "h2d_bytes_per_step": int(h2d / Ke),
                         "d2h_bytes_per_step": int(d2h / Ke), "ms_per_step": e2e_s * 1e3 / Ke, "steps": Ke,
                         "path": "one run_sweep call over %d pressures: msh file -> structure -> H2D -> relaxations -> D2H -> folders "
                                 "written -> gather" % (Ke * S)},
                    gpu_launches=int(launches.item()),
                    submetrics={"tet_updates_per_s": st["tet_updates"] / (st["ms_element"] * 1e-3),
                                "cg_iters_per_s": st["cg_iterations"] / (st["ms_cg"] * 1e-3),
                                "rank0_ms_element": st["ms_element"] / K, "rank0_ms_assembly": st["ms_assembly"] / K,
                                "rank0_ms_cg": st["ms_cg"] / K, "wall_ms_per_step": wall * 1e3 / K,
                                "rank_busy_spread": (ms_max - ms_min) / ms_max if ms_max > 0 else 0.0,
                                "solves_on_rank0": mine, "ms_per_solve": ms_max / K / S * world},
                    roofline=roofline_of(info, st, peaks, peak_kind, args.workload),
                    notes={"l2": "flushed between steps (256 MB write)", "schedule": "shared work counter (dynamic), most expensive first",
                           "nonzero_blocks": info["n_blocks"], "directions": "%d folded to %d" % (info["n_beams"], info["n_dirs"])})
        if world == 1 and not args.no_single:
            try:
                c2 = build_case("cfg2")
                r = single_solve_legs(c2, local, min(K, 5), 3, flush, barrier, world, rank, dist, torch)
                line["single_solve"] = {"workload": c2["text"], "value": 3600.0 / (r["ms_per_step"] * 1e-3), "ms_per_step": r["ms_per_step"],
                                        "e2e_value": 3600.0 / (r["e2e_ms"] * 1e-3), "e2e_ms_per_step": r["e2e_ms"],
                                        "newton_steps": r["newton"], "cg_iterations": r["cg"],
                                        "us_per_cg_iteration": r["st"]["ms_cg"] * 1e3 / max(1, r["st"]["cg_iterations"]),
                                        "ms_element": r["st"]["ms_element"] / min(K, 5), "ms_assembly": r["st"]["ms_assembly"] / min(K, 5),
                                        "ms_cg": r["st"]["ms_cg"] / min(K, 5)}
            except Exception as e:
                line["single_solve"] = {"error": str(e)[:200]}

    if not args.no_hbm_leg:
        try:
            line["roofline_large"] = hbm_regime_leg(local, peaks)
        except Exception as e:   # never lose the main line to the extra leg
            line["roofline_large"] = {"error": str(e)[:200]}
    if not args.no_cpu and world == 1:
        try:
            pool = CpuPool(case, args.cpu_procs)
            try:
                wall, out = pool.wave(wave_pressures(case, args.workload, W, pool.procs))
            finally:
                pool.close()
            line["cpu_baseline"] = {"value": len(out) * 3600.0 / wall, "unit": "sims/hour", "cores": pool.cores, "kind": out[0][3],
                                    "sample": "%d complete solves at once (%d processes x %d threads, %s), %.1f s; Newton steps %s, CG "
                                              "iterations %s" % (len(out), pool.procs, pool.threads, pool.describe(), wall,
                                                                 [o[1] for o in out], [o[2] for o in out])}
        except Exception as e:
            line["cpu_baseline"] = {"error": str(e)[:200]}
    print(json.dumps(line), file=real_stdout)
    real_stdout.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
