#!/bin/bash
cd "$(dirname "$0")/../.."
tag=${1:-r02_f}; out=gpurun_out/$tag; mkdir -p $out
timeout 600 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-extras > $out/bench_cfg4.json 2> $out/bench_cfg4.err
( time timeout 2400 python -m pytest tests -m gpu -q --durations=12 ) > $out/tests.log 2>&1
echo "tests exit $?" >> $out/tests.log
tail -25 $out/tests.log
( time timeout 900 python bench.py --steps 10 --warmup 3 ) > $out/bench_full.json 2> $out/bench_full.err
for f in $out/bench_cfg4.json $out/bench_full.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1]); p=d["phases_ms_per_step"]
    print(sys.argv[1].split('/')[-1], "value %.3fM e2e %.3fM (%.2f ms) pageable %.3fM | setup %.3f prelim %.3f (count %.3f) score %.3f | frac %.3f" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"], d["e2e"].get("pageable",{}).get("value",0)/1e6, p["setup"], p["prelim"], p["prelim_count"], p["score"], d["roofline"]["frac"]), d.get("parity_checked"))
    for k,v in d.get("extra",{}).items(): print("   extra", k, "value %.3fM e2e %.3fM" % (v["value"]/1e6, v["e2e"]["value"]/1e6), v.get("parity_checked"), v.get("index"))
except Exception as e: print(sys.argv[1], "FAILED", e)
PY
done
tail -4 $out/bench_full.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_prelim_wide -s 3 -c 1 -o $out/prof_wide python bench.py --workload cfg4 --spectra 4000 --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $out/ncu_wide.log 2>&1
