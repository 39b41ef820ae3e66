#!/bin/bash
# pageable staging knobs + per-call e2e distribution
cd "$(dirname "$0")/../.."
tag=${1:-r02_stage}; out=gpurun_out/$tag; mkdir -p $out
for cfg in "X=1" "SAGE_B200_STAGE_THREADS=12" "SAGE_B200_STAGE_THREADS=16 SAGE_B200_STAGE_PIECE_KB=1024" "SAGE_B200_STAGE_THREADS=3" "SAGE_B200_STAGE_THREADS=8 SAGE_B200_STAGE_PIECE_KB=512" "SAGE_B200_STAGE_THREADS=24 SAGE_B200_STAGE_PIECE_KB=512"; do
  n=$(echo $cfg | tr '= ' '__')
  env $cfg timeout 300 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > $out/bench_$n.json 2> $out/bench_$n.err
  python - $out/bench_$n.json "$cfg" <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1])
e=d["e2e"]
print(sys.argv[2], "value %.3fM e2e %.3fM (%.3f ms, median call %.3f, max %.3f) pageable %.3fM (%.2f ms)" % (d["value"]/1e6, e["value"]/1e6, e["ms_per_step"], e["ms_per_call_median_rank0"], e["ms_per_call_max_rank0"], e["pageable"]["value"]/1e6, e["pageable"]["ms_per_step"]), d["clocks"])
PY
done
