#!/bin/bash
cd "$(dirname "$0")/../.."
tag=${1:-r02_e}; out=gpurun_out/$tag; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_fullsize.py::test_cfg3_index_matches_oracle_at_size --deselect tests/test_gpu_fullsize.py::test_cfg3_at_size_sample_parity > $out/tests.log 2>&1
echo "tests exit $?" >> $out/tests.log
tail -6 $out/tests.log
timeout 600 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-extras > $out/bench_cfg4.json 2> $out/bench_cfg4.err
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras"
timeout 300 $B > $out/bench_default.json 2> $out/bench_default.err
for pct in 0 10 35; do SAGE_B200_FIRST_CHUNK_PCT=$pct timeout 300 $B > $out/bench_firstchunk$pct.json 2> $out/bench_firstchunk$pct.err; done
for f in $out/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); p=d["phases_ms_per_step"]
    print(sys.argv[1].split('/')[-1], "value %.3fM e2e %.3fM (%.2f ms) pageable %.3fM | setup %.3f prelim %.3f (count %.3f) score %.3f | frac %.3f" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"], d["e2e"].get("pageable",{}).get("value",0)/1e6, p["setup"], p["prelim"], p["prelim_count"], p["score"], d["roofline"]["frac"]), d.get("parity_checked"))
except Exception as e: print(sys.argv[1], "FAILED", e)
PY
done
timeout 900 python bench.py --steps 10 --warmup 3 --extras cfg4,cfg5 > $out/bench_extras.json 2> $out/bench_extras.err
tail -3 $out/bench_extras.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_prelim_wide -s 3 -c 1 -o $out/prof_wide python bench.py --workload cfg4 --spectra 4000 --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $out/ncu_wide.log 2>&1
