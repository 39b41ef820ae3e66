#!/bin/bash
# quick check of a host-pipeline / setup change: GPU tests (without the full-size file), cfg2 bench line, e2e timelines
cd "$(dirname "$0")/../.."
tag=${1:-r02_setup}; out=gpurun_out/$tag; mkdir -p $out
( time timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py ) > $out/tests.log 2>&1
echo "tests exit $?" >> $out/tests.log
tail -4 $out/tests.log
( time timeout 900 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline ) > $out/bench_cfg2.json 2> $out/bench_cfg2.err
SAGE_B200_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2> $out/trace_cfg2.err
python - $out/bench_cfg2.json <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1])
p=d["phases_ms_per_step"]
print("value %.3fM e2e %.3fM (%.2f ms) pageable %.3fM (%.2f ms)" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"], d["e2e"]["pageable"]["value"]/1e6, d["e2e"]["pageable"]["ms_per_step"]), {k: round(v,3) for k,v in p.items()})
PY
grep "chunk base" $out/trace_cfg2.err | tail -8
