#!/bin/bash
# ticketed (persistent) k_prelim_narrow_warp: parity tests on the default build (batch 4), phases for batch 0 / 1 / 4 / 16
cd "$(dirname "$0")/../.."
tag=${1:-r02_ticket}; out=gpurun_out/$tag; mkdir -p $out
( time timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py ) > $out/tests.log 2>&1
echo "tests exit $?" >> $out/tests.log
tail -4 $out/tests.log
for v in main c16 lin8 lin64; do
  lib=sage_b200/lib/ab/$v.so; [ $v = main ] && lib=sage_b200/lib/libsage_b200.so
  SAGE_B200_LIB=$PWD/$lib timeout 600 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > $out/bench_$v.json 2> $out/bench_$v.err
  python - $out/bench_$v.json $v <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1])
    e=d["e2e"]
    print(sys.argv[2], "value %.3fM e2e %.3fM (%.3f ms, in-lib %.3f) pageable %.3fM" % (d["value"]/1e6, e["value"]/1e6, e["ms_per_step"], e["ms_in_library_median_rank0"], e["pageable"]["value"]/1e6), {k: round(v,3) for k,v in d["phases_ms_per_step"].items()})
except Exception as ex:
    print(sys.argv[2], "failed", ex)
PY
done
