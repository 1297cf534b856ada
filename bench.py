#!/usr/bin/env python
"""Benchmark of the relaxation hot path (contract: see the task statement / DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg1|cfg4|cfg5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one complete relaxation (Newton iterations until saenopy's five-energy stop rule) of
the named BASELINE config on each rank's GPU -- a lookup-table "sim".  Ranks share nothing on the
data path (independent load cases, weak scaling); the curve rows are gathered over NCCL.

  value   sims/hour, whole job, device time (CUDA events in the library, max over ranks), inputs
          (mesh structure, targets, start displacements) resident in HBM when the timed region starts
  e2e     the same metric through the reference-facing host call (C-ABI jf_solve_boundarycondition
          behind Solver.solve_boundarycondition): host buffers in, structure build, H2D, relaxation,
          D2H of displacements/forces, radial curve, gather
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (BASELINE config text, length_factor, lf_iso override, material name, pressure)
    "cfg1": ("configs[0]: spherical_inclusion lf=0.2 (lf_iso 0.2), linear K_0=2364, 100 Pa", None, 0.2, "linear2364", 100.0),
    "cfg2": ("configs[1]: spherical_inclusion lf=0.05 (lf_iso 0.1667), collagen12, 100 Pa, single solve", 0.05, None, "collagen12", 100.0),
    "cfg4": ("configs[3]: spherical_inclusion lf=0.03 (lf_iso 0.1), collagen24, 10 kPa", 0.03, None, "collagen24", 10000.0),
    "cfg5": ("configs[4]: spherical_inclusion lf=0.015 (lf_iso 0.05), collagen06, 100 Pa", 0.015, None, "collagen06", 100.0),
}
SOLVER = dict(step=0.0033, max_iter=600, conv_crit=0.01)   # reference defaults, simulation.py:81


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
    r_in, r_out = 100 * 10 ** -6, 10000 * 10 ** -6
    bc = spherical_boundary_conditions(coords, r_in, r_out, pressure)
    movable = np.any(np.isnan(bc["displacement"]), axis=1)
    mat_dict, mat = material_of(matname)
    return dict(text=text, coords=coords, tets=tets.astype(np.int32), movable=movable, ft=np.nan_to_num(bc["forces"]),
                bc=bc, mat=mat, mat_dict=mat_dict, matname=matname, pressure=pressure, r_in=r_in, beams=build_beams(300))


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


def hbm_regime_leg(device, peaks):
    """The same CG kernel family where the matrix cannot be cache resident (BASELINE configs[4] size):
    a structured graded shell mesh (502,968 nodes / 2,989,980 tets), one evaluation, CG launches of exactly
    50 iterations timed with the library's CUDA events."""
    from jointforces_b200 import _lib, materials
    from jointforces_b200.mesh import generate_spherical_inclusion
    from jointforces_b200.simulation import spherical_boundary_conditions
    from jointforces_b200.solver import build_beams
    coords, tets = generate_spherical_inclusion(lf_iso=0.05, method="icoshell")
    bc = spherical_boundary_conditions(coords, 100 * 10 ** -6, 10000 * 10 ** -6, 100.0)
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
            it = ctx.cg(1e-30, 50)
        st, info = ctx.stats(), ctx.info()
    by = cg_bytes_per_iteration(info)
    achieved = by * 50 * reps / (st["ms_cg"] * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": "jf_cg_streamhalo_kernel" if info.get("cg_stream_halo") else "jf_cg_stream_kernel", "workload": "configs[4]-size icoshell mesh, %d nodes / %d tets, "
            "collagen06; matrix %.0f MB" % (info["n_nodes"], info["n_tets"], info["n_blocks"] * 72 / 1e6),
            "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
            "us_per_cg_iteration": st["ms_cg"] * 1e3 / (50 * reps), "algorithmic_bytes_per_launch": by * 50,
            "traffic": None,
            "tet_updates_per_s": st["tet_updates"] / (st["ms_element"] * 1e-3) if st["ms_element"] > 0 else None}


def cpu_sample(case, newton_steps, threads):
    """The oracle's C port on the host cores, bounded to `newton_steps` Newton iterations."""
    from oracle import c_oracle
    t0 = time.time()
    r = c_oracle.relax(case["coords"], case["tets"], case["movable"], case["ft"], np.zeros_like(case["coords"]),
                       case["beams"], case["mat"], SOLVER["step"], newton_steps, SOLVER["conv_crit"], nthreads=threads)
    return time.time() - t0, r


def cpu_full_solve_estimate(r, dt, full_newton, full_cg):
    """Scale a sample of the first Newton iterations to a full solve: the element/assembly part with the number
    of evaluations, the CG part with the number of CG iterations (the early Newton steps need fewer CG iterations
    than the average one, so scaling everything by the Newton count alone would flatter the CPU)."""
    n_s, cg_s = int(r["n_iter"]), int(r["cg_iters"].sum())
    t_el, t_cg = float(r["t_element"]), float(r["t_cg"])
    other = max(0.0, dt - t_el - t_cg)   # pattern build, shape tensors: once per solve
    return other + t_el * (full_newton + 1) / (n_s + 1) + t_cg * full_cg / max(1, cg_s)


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  saenopy itself cannot be
    installed here, so this is the oracle's C port of the same algorithm (kind
    "port"), on all host threads, each step a bounded sample of the workload."""
    if rank != 0:
        return
    case = build_case(args.workload)
    cores = os.cpu_count() or 1
    sample_newton = args.cpu_newton
    full_newton = args.full_newton
    for _ in range(args.warmup and 1):
        cpu_sample(case, 1, cores)
    times, ests = [], []
    for _ in range(args.steps):
        dt, r = cpu_sample(case, sample_newton, cores)
        times.append(dt)
        ests.append(cpu_full_solve_estimate(r, dt, full_newton, args.full_cg))
    est_solve_s = float(np.mean(ests))
    value = 3600.0 / est_solve_s   # the host cores serve the ranks' solves one after another
    sample = ("first %d of %d Newton iterations of one solve (%d nodes, %d tets) with oracle/saeno_oracle.c on %d threads, "
              "%.2f s; element part scaled by evaluations, CG part by CG iterations (%d per solve) -> %.1f s per solve"
              % (sample_newton, full_newton, len(case["coords"]), len(case["tets"]), cores, float(np.mean(times)),
                 args.full_cg, est_solve_s))
    line = {
        "impl": "reference", "metric": "lookup-table sims/hour", "value": value, "unit": "sims/hour", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": case["text"], "n_nodes": int(len(case["coords"])), "n_tets": int(len(case["tets"])),
                   "solver": SOLVER, "step_is": "bounded sample: %d Newton iterations" % sample_newton},
        "cpu_baseline": {"value": value, "unit": "sims/hour", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "sims/hour", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-newton", type=int, default=6, help="Newton iterations in the CPU baseline sample")
    ap.add_argument("--full-newton", type=int, default=86, help="Newton iterations of the full solve (cfg2: 86)")
    ap.add_argument("--full-cg", type=int, default=68800, help="CG iterations of the full solve (cfg2: ~68,800)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-hbm-leg", action="store_true", help="skip the large-mesh (HBM-bound) CG roofline leg")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from jointforces_b200 import _lib
    from jointforces_b200.simulation import radial_median_curve, lookup_bins
    from jointforces_b200.sweep import gather_table

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
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    t_wall0 = time.time()
    for _ in range(K):
        relrec, cgit, n_iter = device_step()
    barrier()
    wall = time.time() - t_wall0
    clocks = sampler.stop()
    st = ctx.stats()
    ms_dev = torch.tensor([st["ms_relax"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms_dev, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms_dev.item()) / K
    value = world * 3600.0 / (ms_per_step * 1e-3)

    # ---- end to end through the host-facing call -----------------------------------------------
    def e2e_step():
        out = _lib.solve_boundarycondition(case["coords"], case["tets"], case["movable"], case["ft"],
                                           np.zeros_like(case["coords"]), case["mat"], SOLVER["step"], SOLVER["max_iter"],
                                           SOLVER["conv_crit"], beams=case["beams"], device=local)
        curve = radial_median_curve(case["coords"], out["U"], case["r_in"], x_edges)
        gather_table([rank], [case["pressure"]], curve[None, :], world, len(curve))
        return out

    e2e_step()
    barrier()
    t0 = time.time()
    for _ in range(K):
        flush.fill_(1)
        out = e2e_step()
    barrier()
    e2e_ms = torch.tensor([(time.time() - t0) * 1e3 / K], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * 3600.0 / (float(e2e_ms.item()) * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_kind = measured_peaks()
    cg_iters = st["cg_iterations"]
    by_iter = cg_bytes_per_iteration(info)
    achieved = by_iter * cg_iters / (st["ms_cg"] * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "cg_traffic_%s.json" % args.workload)
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    matrix_mb = info["n_blocks"] * 72 / 1e6
    line = {
        "metric": "lookup-table sims/hour", "value": value, "unit": "sims/hour", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": case["text"], "n_nodes": info["n_nodes"], "n_free": info["n_free"], "n_tets": info["n_tets"],
                   "nonzero_blocks": info["n_blocks"], "directions": "%d folded to %d" % (info["n_beams"], info["n_dirs"]),
                   "solver": SOLVER, "newton_steps": int(n_iter[0]), "cg_iterations_per_solve": int(cgit[0].sum()),
                   "l2": "flushed between steps (256 MB write)", "parallelism": "independent solves, 1 per GPU"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "sims/hour", "h2d_bytes_per_step": int(out["stats"]["bytes_h2d"]),
                "d2h_bytes_per_step": int(out["stats"]["bytes_d2h"]), "ms_per_step": float(e2e_ms.item())},
        "gpu_launches": int(st["kernel_launches"]),
        "submetrics": {
            "tet_updates_per_s": st["tet_updates"] / (st["ms_element"] * 1e-3),
            "cg_iters_per_s": cg_iters / (st["ms_cg"] * 1e-3),
            "ms_element": st["ms_element"] / K, "ms_assembly": st["ms_assembly"] / K, "ms_cg": st["ms_cg"] / K,
            "wall_ms_per_step": wall * 1e3 / K,
        },
        "roofline": {"bound": "hbm", "kernel": "jf_cg_fused_kernel" if info.get("cg_fused") == 1 else "jf_cg_fused_stream_kernel" if info.get("cg_fused") == 2 else
                     "jf_cg_regres_kernel" if info["cg_resident"] == 2 else
                     ("jf_cg_resident_kernel" if info["cg_resident"] == 1 else
                      ("jf_cg_streamhalo_kernel" if info.get("cg_stream_halo") else "jf_cg_stream_kernel")), "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_kind": "of " + peak_kind,
                     "algorithmic_bytes_per_launch": by_iter * cg_iters / max(1, st["cg_launches"]),
                     "note": ("matrix %.0f MB %s the 126 MB L2; " % (matrix_mb, "fits in" if matrix_mb < 100 else "exceeds")) +
                             ("it is staged once per CG solve into registers/shared memory, so `achieved` is a rate of "
                              "algorithmic bytes (latency/barrier-bound regime), not DRAM traffic; `traffic` is the measured "
                              "DRAM bytes of one launch; see roofline_large for the HBM-bound regime" if info["cg_resident"] else "")},
    }
    if not args.no_hbm_leg:
        try:
            line["roofline_large"] = hbm_regime_leg(local, peaks)
        except Exception as e:   # never lose the main line to the extra leg
            line["roofline_large"] = {"error": str(e)[:200]}
    if not args.no_cpu:
        cores = os.cpu_count() or 1
        dt, r = cpu_sample(case, args.cpu_newton, cores)
        full, full_cg = int(n_iter[0]), int(cgit[0].sum())
        est = cpu_full_solve_estimate(r, dt, full, full_cg)
        line["cpu_baseline"] = {"value": 3600.0 / est, "unit": "sims/hour", "cores": cores, "kind": "port",
                                "sample": "first %d of %d Newton iterations of the same solve with oracle/saeno_oracle.c on %d "
                                          "threads, %.1f s; element part scaled by evaluations, CG part by CG iterations "
                                          "(%d per solve) -> %.1f s per solve" % (args.cpu_newton, full, cores, dt, full_cg, est)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
