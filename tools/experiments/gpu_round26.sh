#!/bin/bash
cd "$(dirname "$0")/../.."
tag=${1:-r02_cfg3blk}; out=gpurun_out/$tag; mkdir -p $out
for cfg in "SAGE_B200_NARROW_BLOCK=1024" "SAGE_B200_NARROW_BLOCK=2048" "SAGE_B200_NARROW_BLOCK=512"; do
  n=$(echo $cfg | tr '= ' '__')
  env $cfg timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-extras --no-cpu-baseline > $out/bench_cfg3_$n.json 2> $out/bench_cfg3_$n.err
  python - $out/bench_cfg3_$n.json "$cfg" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1])
    e=d["e2e"]
    print(sys.argv[2], "value %.3fM e2e %.3fM (%.3f ms)" % (d["value"]/1e6, e["value"]/1e6, e["ms_per_step"]), {k: round(v,3) for k,v in d["phases_ms_per_step"].items() if k in ("setup","prelim","prelim_count","score")}, d["work_per_step"]["queries"])
except Exception as ex:
    print(sys.argv[2], "failed", ex)
PY
done
