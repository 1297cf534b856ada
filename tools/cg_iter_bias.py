"""Signed CG-iteration difference per Newton step between the kernels and the committed oracle solve of config 2:
python tools/cg_iter_bias.py   (a systematic bias would hide behind |d| <= 12)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import build_case, SOLVER
from jointforces_b200 import _lib

g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "oracle_full_cfg2_full.npz"))
case = build_case("cfg2")
runs = {}
for label, env in (("fused+fixed-point", {}), ("fused+slot-sums", {"JF_CG_SLOT_SUMS": "1"}), ("classic-regres", {"JF_CG_CLASSIC": "1"}),
                   ("classic-stream", {"JF_CG_CLASSIC": "1", "FLAGS": "4"})):
    flags = int(env.pop("FLAGS", "0"))
    for k, v in env.items():
        os.environ[k] = v
    with _lib.Context(case["coords"], case["tets"], case["movable"], n_sys=1, beams=case["beams"], flags=flags) as ctx:
        ctx.set_material(*case["mat"]); ctx.set_targets(case["ft"]); ctx.set_displacements(np.zeros_like(case["coords"]))
        relrec, cgit, n = ctx.relax(SOLVER["step"], SOLVER["max_iter"], SOLVER["conv_crit"])
        U = ctx.displacements()
        info = ctx.info()
    for k in env:
        del os.environ[k]
    runs[label] = (cgit[0].astype(np.int64), U)
    d = cgit[0].astype(np.int64)[: len(g["cg_iters"])] - g["cg_iters"][: len(cgit[0])]
    print("%-18s resident %d fused %d newton %d cg %d vs oracle %d : signed mean %+.2f/step, std %.2f, max|d| %d, err U %.2e"
          % (label, info["cg_resident"], info["cg_fused"], n[0], cgit[0].sum(), g["cg_iters"].sum(), d.mean(), d.std(), np.abs(d).max(),
             np.linalg.norm(U - g["U_sub"]) / np.linalg.norm(g["U_sub"])))
    print("   first 20 d:", d[:20].tolist())
base = runs["classic-regres"][0]
for label, (c, U) in runs.items():
    m = min(len(c), len(base))
    print("%-18s vs classic-regres: signed mean %+.2f" % (label, (c[:m] - base[:m]).mean()))
