#!/bin/bash
# WALK_SOLO x COOP_GROUP sweep of the block probe (cfg2 phases)
cd "$(dirname "$0")/../.."
tag=${1:-r02_coop2}; out=gpurun_out/$tag; mkdir -p $out
for v in s16g32 s32g32 s4g8 s8g8 s8g4 s12g8; do
  lib=sage_b200/lib/ab/$v.so
  SAGE_B200_LIB=$PWD/$lib timeout 600 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > $out/bench_$v.json 2> $out/bench_$v.err
  python - $out/bench_$v.json $v <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1])
    e=d["e2e"]
    print(sys.argv[2], "value %.3fM e2e %.3fM (%.3f ms, in-lib %.3f) psms %d" % (d["value"]/1e6, e["value"]/1e6, e["ms_per_step"], e["ms_in_library_median_rank0"], d["psms_per_step_rank0"]), {k: round(v,3) for k,v in d["phases_ms_per_step"].items() if k in ("setup","prelim","prelim_count","score")})
except Exception as ex:
    print(sys.argv[2], "failed", ex)
PY
done
