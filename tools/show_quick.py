"""Summarise a tools/gpu_quick.sh run: python tools/show_quick.py TAG"""
import glob, json, re, sys
import numpy as np
tag = sys.argv[1]
for f in sorted(glob.glob("gpurun_out/%s_bench*.json" % tag)):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print("%-40s value %.0f  ms %.2f  e2e %.2f  cg %.2f  us/it %.3f" % (f.split("/")[-1], d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"],
              d["submetrics"]["ms_cg"], 1e6 / d["submetrics"]["cg_iters_per_s"]))
    except Exception as e:
        print(f, "unreadable:", e, open(f.replace(".json", ".err")).read()[-400:])
for f in sorted(glob.glob("gpurun_out/%s_cgtiming*.log" % tag)):
    rows, names = [], None
    for l in open(f):
        if l.startswith("cgt cta"):
            names = re.findall(r"([a-z_]+) [-+]?\d", l)
            rows.append([float(x) for x in re.findall(r"(?<![a-z_])[-+]?\d+\.?\d*", l)])
    if not rows:
        print(f, "no timing rows"); continue
    a = np.array(rows)
    b = a[a[:, 3] == a[0, 3]]
    i0 = names.index("top")
    print("%-40s total %.0f | " % (f.split("/")[-1], b[:, i0:-1].sum(1).mean()) + " ".join("%s %.0f" % (names[k], b[:, k].mean()) for k in range(i0, len(names))))
for l in open("gpurun_out/%s_tests.log" % tag).read().splitlines()[-3:]:
    print(l)
