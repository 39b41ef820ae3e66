#!/bin/bash
cd "$(dirname "$0")/../.."
tag=${1:-r02_j}; out=gpurun_out/$tag; mkdir -p $out
B="python bench.py --workload cfg4 --steps 5 --warmup 3 --no-extras"
timeout 600 $B > $out/bench_cfg4_default.json 2> $out/bench_cfg4_default.err
for v in sage_b200/lib/variants/libsage_b200_wide*.so; do
  n=$(basename $v .so); n=${n#libsage_b200_}
  SAGE_B200_LIB=$PWD/$v timeout 600 $B > $out/bench_cfg4_$n.json 2> $out/bench_cfg4_$n.err
done
for f in $out/bench_cfg4_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1]); p=d["phases_ms_per_step"]
    print(sys.argv[1].split('/')[-1], "value %.3fM e2e %.3fM | prelim %.3f | frac %.3f overflows %d" % (d["value"]/1e6, d["e2e"]["value"]/1e6, p["prelim"], d["roofline"]["frac"], d["work_per_step"]["wide_overflows"]), d.get("parity_checked"))
except Exception as e: print(sys.argv[1], "FAILED", e); print(open(sys.argv[1].replace('.json','.err')).read()[-600:])
PY
done
