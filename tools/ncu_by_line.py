#!/usr/bin/env python
"""Per-source-line instruction / stall shares of one kernel from an `ncu --set full --import-source on` report.

    python tools/ncu_by_line.py <report.ncu-rep> <library.so> <kernel-name-substring> [top_n]

Joins the SASS page of the report (`ncu --page source --csv`: per-instruction executed counts, thread counts, stall samples) with the
line table of the same kernel in the library (`cuobjdump -xelf` + `nvdisasm -g`): the n-th instruction of the report is the n-th
instruction of the disassembly. Prints, per source line, its share of executed warp instructions, of stall samples, and the average
number of active threads per executed instruction ("eff", 32 = no divergence).
"""
import csv
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(so, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
    cubins = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")]
    out = []
    for cb in cubins:
        txt = subprocess.check_output(["nvdisasm", "-g", "-c", cb]).decode(errors="replace")
        cur_fn, cur_line, active = None, ("?", 0), False
        for ln in txt.splitlines():
            m = re.match(r"\s*\.text\.(\S+):", ln)
            if m:
                cur_fn = m.group(1)
                active = kernel in cur_fn
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m and active:
                out.append((int(m.group(1), 16), cur_line, m.group(2).strip()))
        if out:
            break
    return out


def main():
    rep, so, kernel = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kernel], stderr=subprocess.DEVNULL).decode()
    rows = list(csv.reader(raw.splitlines()))
    # a report with several kernels prints one block per launch: "Kernel Name" line, column header ("Address", ...), instructions
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] or [0]
    blocks = [(rows[a][1] if len(rows[a]) > 1 else "", rows[a:b]) for a, b in zip(starts, starts[1:] + [len(rows)])]
    name, blk = next(((n, b) for n, b in blocks if kernel in n.split("(")[0]), blocks[0])
    hi = next(i for i, r in enumerate(blk) if r and r[0] == "Address")
    hdr = blk[hi]
    col = {h: i for i, h in enumerate(hdr)}
    inst = [r for r in blk[hi + 1:] if len(r) == len(hdr) and r[0] != "Address"]
    sass = sass_lines(so, os.environ.get("SASS_NAME") or kernel)   # SASS_NAME: mangled-name substring when several instantiations share the name
    if len(sass) != len(inst):
        print(f"# warning: {len(inst)} instructions in the report vs {len(sass)} in the disassembly (different build?)", file=sys.stderr)
    agg = {}
    tot_i = tot_s = tot_t = 0
    for k, r in enumerate(inst):
        line = sass[k][1] if k < len(sass) else ("?", 0)
        ie, te, ss = float(r[col["Instructions Executed"]]), float(r[col["Thread Instructions Executed"]]), float(r[col["# Samples"]])
        a = agg.setdefault(line, [0.0, 0.0, 0.0])
        a[0] += ie; a[1] += te; a[2] += ss
        tot_i += ie; tot_t += te; tot_s += ss
    srcs = {}
    def text(f, n):
        if f not in srcs:
            p = os.path.join(os.environ.get("SAGE_B200_SRC_DIR") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "sage_b200", "csrc"), f)
            srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
        return srcs[f][n - 1].strip()[:110] if 0 < n <= len(srcs[f]) else ""
    print(f"# {kernel}: total warp inst {tot_i:.0f} samples {tot_s:.0f} thread-eff {tot_t / max(tot_i, 1):.2f}")
    by = 2 if os.environ.get("SORT_BY_STALL") else 0   # SORT_BY_STALL=1: rank lines by stall samples instead of executed instructions
    for line, a in sorted(agg.items(), key=lambda kv: -kv[1][by])[:top]:
        print(f"{line[0][:14]:14s}:{line[1]:4d} inst {100 * a[0] / tot_i:5.1f}% stall {100 * a[2] / max(tot_s, 1):5.1f}% eff {a[1] / max(a[0], 1):4.1f} | {text(*line)}")


if __name__ == "__main__":
    main()
