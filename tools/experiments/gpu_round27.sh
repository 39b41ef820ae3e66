#!/bin/bash
# automatic block size: parity tests + full bench with extras
cd "$(dirname "$0")/../.."
tag=${1:-r02_auto}; out=gpurun_out/$tag; mkdir -p $out
( time timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py ) > $out/tests.log 2>&1
echo "tests exit $?" >> $out/tests.log
tail -4 $out/tests.log
( time timeout 900 python bench.py --steps 20 --warmup 3 ) > $out/bench_full.json 2> $out/bench_full.err
python - $out/bench_full.json <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1])
e=d["e2e"]
print("value %.3fM e2e %.3fM (%.3f ms, in-lib %.3f) pageable %s" % (d["value"]/1e6, e["value"]/1e6, e["ms_per_step"], e["ms_in_library_median_rank0"], e.get("pageable",{}).get("value")), {k: round(v,3) for k,v in d["phases_ms_per_step"].items() if k in ("setup","prelim","prelim_count","score")}, d.get("parity_checked",{}).get("psms_identical_to_oracle"), d["index"]["hbm_bytes"], d["roofline"]["frac"], d.get("cpu_baseline",{}).get("value"))
for k,v in d.get("extra",{}).items(): print("   extra", k, ("value %.3fM e2e %.3fM" % (v["value"]/1e6, v["e2e"]["value"]/1e6)) if "value" in v else v, {a: round(b,3) for a,b in v.get("phases_ms_per_step",{}).items() if a in ("prelim","prelim_count","score")}, v.get("parity_checked",{}).get("psms_identical_to_oracle"), v["roofline"]["frac"])
PY
