#!/bin/bash
# One gpurun call: fast GPU parity subset, then A/B bench lines of cfg2, then ncu of k_score. Output under gpurun_out/<tag>/.
cd "$(dirname "$0")/../.."
tag=${1:-r02_a}; out=gpurun_out/$tag; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
timeout 900 python -m pytest tests/test_glibc_log.py tests/test_gpu_parity.py tests/test_gpu_random.py tests/test_gpu_process.py -m gpu -x -q > $out/tests.log 2>&1
echo "tests exit $?" >> $out/tests.log
tail -5 $out/tests.log
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline"
timeout 300 $B > $out/bench_default.json 2> $out/bench_default.err
for v in 10 12; do
  for t in 1024 2048; do
    if [ $v = 12 ] && [ $t = 2048 ]; then continue; fi
    SAGE_B200_LIB=$PWD/sage_b200/lib/variants/libsage_b200_ctas$v.so SAGE_B200_SCORE_TILE=$t timeout 300 $B > $out/bench_ctas${v}_tile$t.json 2> $out/bench_ctas${v}_tile$t.err
  done
done
SAGE_B200_SCORE_TILE=1024 timeout 300 $B > $out/bench_ctas8_tile1024.json 2> $out/bench_ctas8_tile1024.err
SAGE_B200_SCORE_TILE=512 timeout 300 $B > $out/bench_ctas8_tile512.json 2> $out/bench_ctas8_tile512.err
for f in $out/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); p=d["phases_ms_per_step"]
    print(sys.argv[1].split('/')[-1], "value %.2fM e2e %.2fM | setup %.3f prelim %.3f (count %.3f) score %.3f" % (d["value"]/1e6, d["e2e"]["value"]/1e6, p["setup"], p["prelim"], p["prelim_count"], p["score"]))
except Exception as e: print(sys.argv[1], "FAILED", e)
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_score -s 6 -c 1 -o $out/prof_kscore python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/ncu_kscore.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/launches.log 2>&1
ls -la $out
