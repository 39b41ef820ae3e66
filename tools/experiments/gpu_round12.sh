#!/bin/bash
# A/B of k_prelim_narrow_warp variants (flattened probe loop): parity tests on the main build, then phases of every variant.
cd "$(dirname "$0")/../.."
tag=${1:-r02_flat}; out=gpurun_out/$tag; mkdir -p $out
( time timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_random.py -m gpu -q -x ) > $out/tests.log 2>&1
echo "tests exit $?" >> $out/tests.log
tail -4 $out/tests.log
for v in main old r2 r1c16 r2c16; do
  lib=sage_b200/lib/ab/$v.so; [ $v = main ] && lib=sage_b200/lib/libsage_b200.so
  [ -f $lib ] || continue
  SAGE_B200_LIB=$PWD/$lib timeout 600 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > $out/bench_$v.json 2> $out/bench_$v.err
  python - $out/bench_$v.json $v <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1])
    p=d["phases_ms_per_step"]
    print(sys.argv[2], "value %.3fM e2e %.3fM (%.2f ms)" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"]), {k: round(v,3) for k,v in p.items()}, d.get("parity_checked",{}).get("psms_identical_to_oracle"))
except Exception as e:
    print(sys.argv[2], "failed", e)
PY
done
