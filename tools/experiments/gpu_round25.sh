#!/bin/bash
# block index in the CTA counting kernel: full GPU test suite (incl. cfg3 at size), cfg3 with / without it
cd "$(dirname "$0")/../.."
tag=${1:-r02_cta}; out=gpurun_out/$tag; mkdir -p $out
( time timeout 2400 python -m pytest tests -m gpu -q ) > $out/tests.log 2>&1
echo "tests exit $?" >> $out/tests.log
tail -4 $out/tests.log
for cfg in "SAGE_B200_NARROW_CTA=1" "SAGE_B200_NARROW_CTA=0"; do
  n=$(echo $cfg | tr '= ' '__')
  env $cfg timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-extras --no-cpu-baseline > $out/bench_cfg3_$n.json 2> $out/bench_cfg3_$n.err
  python - $out/bench_cfg3_$n.json "$cfg" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1])
    e=d["e2e"]
    print(sys.argv[2], "value %.3fM e2e %.3fM (%.3f ms)" % (d["value"]/1e6, e["value"]/1e6, e["ms_per_step"]), {k: round(v,3) for k,v in d["phases_ms_per_step"].items() if k in ("setup","prelim","prelim_count","score")}, d["index"]["hbm_bytes"])
except Exception as ex:
    print(sys.argv[2], "failed", ex)
PY
done
