#!/bin/bash
# split scoring: GPU tests (without the full-size file), cfg2 bench split vs fused
cd "$(dirname "$0")/../.."
tag=${1:-r02_split}; out=gpurun_out/$tag; mkdir -p $out
( time timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py ) > $out/tests.log 2>&1
echo "tests exit $?" >> $out/tests.log
tail -25 $out/tests.log | cut -c1-250
for cfg in "SAGE_B200_SCORE_SPLIT=1" "SAGE_B200_SCORE_SPLIT=0"; do
  n=$(echo $cfg | tr '= ' '__')
  env $cfg timeout 300 python bench.py --steps 20 --warmup 3 --no-extras > $out/bench_$n.json 2> $out/bench_$n.err
  python - $out/bench_$n.json "$cfg" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1])
    e=d["e2e"]
    print(sys.argv[2], "value %.3fM e2e %.3fM (%.3f ms, in-lib %.3f) pageable %.3fM" % (d["value"]/1e6, e["value"]/1e6, e["ms_per_step"], e["ms_in_library_median_rank0"], e["pageable"]["value"]/1e6), {k: round(v,3) for k,v in d["phases_ms_per_step"].items() if k in ("setup","prelim","prelim_count","score")}, d.get("parity_checked",{}).get("psms_identical_to_oracle"))
except Exception as ex:
    print(sys.argv[2], "failed", ex); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
