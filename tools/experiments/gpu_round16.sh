#!/bin/bash
# masses copy in parts + staging pool: GPU tests, then cfg2 bench lines for 1..4 parts
cd "$(dirname "$0")/../.."
tag=${1:-r02_parts}; out=gpurun_out/$tag; mkdir -p $out
( time timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py ) > $out/tests.log 2>&1
echo "tests exit $?" >> $out/tests.log
tail -4 $out/tests.log
for cfg in "SAGE_B200_MASS_PARTS=2" "SAGE_B200_MASS_PARTS=1" "SAGE_B200_MASS_PARTS=3" "SAGE_B200_MASS_PARTS=4"; do
  n=$(echo $cfg | tr '= ' '__')
  env $cfg timeout 300 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > $out/bench_$n.json 2> $out/bench_$n.err
  python - $out/bench_$n.json "$cfg" <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1])
e=d["e2e"]
print(sys.argv[2], "value %.3fM e2e %.3fM (%.3f ms, median call %.3f, max %.3f) pageable %.3fM (%.2f ms)" % (d["value"]/1e6, e["value"]/1e6, e["ms_per_step"], e["ms_per_call_median_rank0"], e["ms_per_call_max_rank0"], e["pageable"]["value"]/1e6, e["pageable"]["ms_per_step"]), d["clocks"])
PY
  env $cfg SAGE_B200_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2> $out/trace_$n.err
  grep "chunk base" $out/trace_$n.err | sed -n '5,6p;10,11p'
done
