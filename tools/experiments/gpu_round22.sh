#!/bin/bash
# narrow block index as default: GPU tests, full bench line (extras incl. cfg3 / cfg5 with parity), vector-walk variant on cfg2, old path on cfg3/cfg5 for comparison
cd "$(dirname "$0")/../.."
tag=${1:-r02_nblk3}; out=gpurun_out/$tag; mkdir -p $out
( time timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py ) > $out/tests.log 2>&1
echo "tests exit $?" >> $out/tests.log
tail -4 $out/tests.log
( time timeout 900 python bench.py --steps 20 --warmup 3 ) > $out/bench_full.json 2> $out/bench_full.err
SAGE_B200_LIB=$PWD/sage_b200/lib/ab/vec.so timeout 300 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > $out/bench_vec.json 2> $out/bench_vec.err
SAGE_B200_NARROW_INDEX=0 timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-extras --no-cpu-baseline > $out/bench_cfg3.json 2> $out/bench_cfg3.err
for f in $out/bench_full.json $out/bench_vec.json $out/bench_cfg3.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1])
    e=d["e2e"]
    print(sys.argv[1].split('/')[-1], "value %.3fM e2e %.3fM (%.3f ms, in-lib %.3f) pageable %s" % (d["value"]/1e6, e["value"]/1e6, e["ms_per_step"], e["ms_in_library_median_rank0"], e.get("pageable",{}).get("value")), {k: round(v,3) for k,v in d["phases_ms_per_step"].items() if k in ("setup","prelim","prelim_count","score")}, d.get("parity_checked",{}).get("psms_identical_to_oracle"), d["index"]["hbm_bytes"])
    for k,v in d.get("extra",{}).items(): print("   extra", k, ("value %.3fM e2e %.3fM" % (v["value"]/1e6, v["e2e"]["value"]/1e6)) if "value" in v else v, {a: round(b,3) for a,b in v.get("phases_ms_per_step",{}).items() if a in ("prelim","prelim_count","score")}, v.get("parity_checked",{}).get("psms_identical_to_oracle"))
except Exception as ex:
    print(sys.argv[1], "failed", ex)
PY
done
