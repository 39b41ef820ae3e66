#!/bin/bash
cd "$(dirname "$0")/../.."
tag=${1:-r02_c}; out=gpurun_out/$tag; mkdir -p $out
python tools/debug_seed.py 23 > $out/seed23.log 2>&1
timeout 900 python -m pytest tests/test_glibc_log.py tests/test_gpu_parity.py tests/test_gpu_random.py tests/test_gpu_process.py -m gpu -q > $out/tests.log 2>&1
echo "tests exit $?" >> $out/tests.log
tail -5 $out/tests.log
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline"
timeout 300 $B > $out/bench_default.json 2> $out/bench_default.err
for v in sage_b200/lib/variants/*.so; do
  n=$(basename $v .so); n=${n#libsage_b200_}
  SAGE_B200_LIB=$PWD/$v timeout 300 $B > $out/bench_$n.json 2> $out/bench_$n.err
done
SAGE_B200_SCORE_FAST=0 timeout 300 $B > $out/bench_generic.json 2> $out/bench_generic.err
for f in $out/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); p=d["phases_ms_per_step"]
    print(sys.argv[1].split('/')[-1], "value %.2fM e2e %.2fM | setup %.3f prelim %.3f (count %.3f) score %.3f" % (d["value"]/1e6, d["e2e"]["value"]/1e6, p["setup"], p["prelim"], p["prelim_count"], p["score"]))
except Exception as e: print(sys.argv[1], "FAILED", e)
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_score -s 6 -c 1 -o $out/prof_kscore python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/ncu_kscore.log 2>&1
