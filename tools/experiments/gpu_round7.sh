#!/bin/bash
cd "$(dirname "$0")/../.."
tag=${1:-r02_g}; out=gpurun_out/$tag; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_random.py tests/test_gpu_fullsize.py::test_cfg4_sample_parity tests/test_gpu_fullsize.py::test_cfg2_sample_parity_and_properties -m gpu -q > $out/tests.log 2>&1
echo "tests exit $?" >> $out/tests.log
tail -4 $out/tests.log
timeout 600 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-extras > $out/bench_cfg4.json 2> $out/bench_cfg4.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > $out/bench_cfg2.json 2> $out/bench_cfg2.err
for f in $out/bench_cfg4.json $out/bench_cfg2.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1]); p=d["phases_ms_per_step"]
    print(sys.argv[1].split('/')[-1], "value %.3fM e2e %.3fM (%.2f ms) pageable %.3fM | setup %.3f prelim %.3f (count %.3f) score %.3f | frac %.3f" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"], d["e2e"].get("pageable",{}).get("value",0)/1e6, p["setup"], p["prelim"], p["prelim_count"], p["score"], d["roofline"]["frac"]), d.get("parity_checked"))
except Exception as e: print(sys.argv[1], "FAILED", e)
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_prelim_wide -s 3 -c 1 -o $out/prof_wide python bench.py --workload cfg4 --spectra 4000 --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $out/ncu_wide.log 2>&1
