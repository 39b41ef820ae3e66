#!/usr/bin/env python
"""Turns the scratch output of tools/gpu_final_profile.sh (gpurun_out/<tag>/) into the committed summaries under profiles/ (names r02_final_*):
bench JSON lines, ncu launch list + per-kernel summary, key metrics of the `ncu --set full` captures, per-source-line tables, traffic.json.

    python tools/collect_profiles.py gpurun_out/r02_final3 sage_b200/lib/libsage_b200.so [tag]
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
src, lib = sys.argv[1], sys.argv[2]
P = os.path.join(ROOT, "profiles")
tag = sys.argv[3] if len(sys.argv) > 3 else "r02_final3"


def json_line(path):
    for ln in reversed(open(path).read().strip().splitlines()):
        if ln.startswith("{"):
            return json.loads(ln)
    raise RuntimeError("no JSON line in " + path)


for name in ("bench_cfg2_full", "bench_reference_arm", "bench_cfg4", "bench_cfg5", "bench_cfg3", "bench_cfg2_page_index", "bench_cfg2_fused"):
    f = os.path.join(src, name + ".json")
    if os.path.exists(f):
        try:
            json.dump(json_line(f), open(os.path.join(P, f"{tag}_{name}.json"), "w"), indent=1)
        except Exception as e:
            print("skip", name, e)

# launch list + summary
lf = os.path.join(src, "launches_cfg2.csv")
if os.path.exists(lf):
    rows = list(csv.reader(open(lf)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[hi]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    with open(os.path.join(P, f"{tag}_launches_cfg2.csv"), "w") as o:
        o.write("# ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 80: python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras\nkernel,duration_us\n")
        agg = {}
        for r in rows[hi + 1:]:
            if len(r) > vi:
                k = r[ki].split("(")[0][-60:]
                v = float(r[vi]) / 1000.0
                o.write(f"{k},{v:.1f}\n")
                a = agg.setdefault(k, [0, 0.0])
                a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(P, f"{tag}_launch_summary_cfg2.csv"), "w") as o:
        o.write("kernel,launches,total_us,share\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            o.write(f"{k},{a[0]},{a[1]:.1f},{a[1] / tot:.3f}\n")

traffic = {"_sources_sha": None, "_source": f"ncu --set full captures of tools/gpu_final_profile.sh ({tag}): dram__bytes_read.sum + dram__bytes_write.sum per launch"}
import bench
traffic["_sources_sha"] = bench.kernel_sources_sha()
for rep, wl, kernels in (("prof_cfg2.ncu-rep", "cfg2", ["k_score", "k_prelim_narrow_warp", "k_replay", "k_setup_queries", "k_fold", "k_features", "k_rows"]), ("prof_cfg4.ncu-rep", "cfg4", ["k_prelim_wide", "k_wide_account"])):
    rp = os.path.join(src, rep)
    if not os.path.exists(rp):
        continue
    summ = json.loads(subprocess.check_output([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rp]).decode())
    json.dump(summ, open(os.path.join(P, f"{tag}_ncu_full_{wl}.json"), "w"), indent=1)
    traffic[wl] = {}
    for e in summ:
        for k in kernels:
            if k in e["Kernel Name"].split("(")[0]:
                def num(x):
                    v, u = x.split()[0], x.split()[1] if len(x.split()) > 1 else ""
                    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                    return float(v) * m
                tb = num(e["dram__bytes_read.sum"]) + num(e["dram__bytes_write.sum"])
                if tb == tb:   # ncu reports nan for a launch it could not attribute DRAM counters to: keep the kernel out of traffic.json
                    traffic[wl][k] = int(tb)
    for k in kernels[:3] if wl == "cfg2" else kernels[:1]:
        try:
            env = dict(os.environ)
            if k == "k_prelim_narrow_warp":
                env["SASS_NAME"] = "k_prelim_narrow_warpILb1"   # the block-index instantiation is the one the timed steps run
            out = subprocess.check_output([sys.executable, os.path.join(ROOT, "tools", "ncu_by_line.py"), rp, lib, k, "45"], stderr=subprocess.DEVNULL, env=env).decode()
            open(os.path.join(P, f"{tag}_{k}_by_source_line.txt"), "w").write(out)
        except Exception as e:
            print("by-line failed", k, e)
if "cfg2" in traffic and "k_score" in traffic["cfg2"]:   # bench.py times the four scoring kernels as one phase under the name k_score
    parts = {k: traffic["cfg2"].get(k, 0) for k in ("k_score", "k_fold", "k_features", "k_rows")}
    traffic["cfg2"]["k_score_match_only"] = parts["k_score"]
    traffic["cfg2"]["k_score"] = int(sum(parts.values()))
    traffic["_note_k_score"] = "cfg2.k_score = k_score<true> + k_fold + k_features + k_rows (the split scoring phase); k_score_match_only = the first of them"
json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
for name in ("sanitizer_memcheck.txt", "tests.log", "trace_cfg2.err", "smi.txt", "phase_cycles_cfg2.txt"):
    f = os.path.join(src, name)
    if os.path.exists(f):
        txt = open(f).read()
        open(os.path.join(P, f"{tag}_{name.replace('.err', '.txt').replace('.log', '.txt')}"), "w").write(txt[-6000:])
print("done")
