#!/bin/bash
# Final 1-GPU measurement + profiling call of the round: bench lines, reference arm, ncu launch list, ncu --set full of every hot kernel, sanitizer.
cd "$(dirname "$0")/.."
tag=${1:-r02_final4}; out=gpurun_out/$tag; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $out/smi.txt 2>&1
( time timeout 2400 python -m pytest tests -m gpu -q ) > $out/tests.log 2>&1
echo "tests exit $?" >> $out/tests.log
tail -4 $out/tests.log
SAGE_B200_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2> $out/trace_cfg2.err
( time timeout 900 python bench.py --steps 20 --warmup 3 ) > $out/bench_cfg2_full.json 2> $out/bench_cfg2_full.err
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 3 ) > $out/bench_reference_arm.json 2> $out/bench_reference_arm.err
SAGE_B200_SCORE_SPLIT=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > $out/bench_cfg2_fused.json 2> $out/bench_cfg2_fused.err
[ -f sage_b200/lib/ab/phase.so ] && SAGE_B200_LIB=$PWD/sage_b200/lib/ab/phase.so timeout 600 python tools/phase_cycles.py cfg2 > $out/phase_cycles_cfg2.txt 2>&1
for f in $out/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1])
    p=d.get("phases_ms_per_step",{})
    print(sys.argv[1].split('/')[-1], "value %.3fM e2e %.3fM" % (d["value"]/1e6, d["e2e"]["value"]/1e6), {k: round(v,3) for k,v in p.items() if k in ("setup","prelim","prelim_count","score")}, "frac", round(d.get("roofline",{}).get("frac",0),3), d.get("parity_checked",{}).get("psms_identical_to_oracle"), d.get("cpu_baseline",{}).get("value"))
    for k,v in d.get("extra",{}).items(): print("   extra", k, ("value %.3fM e2e %.3fM" % (v["value"]/1e6, v["e2e"]["value"]/1e6)) if "value" in v else v)
except Exception as e: print(sys.argv[1], "FAILED", e)
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 80 --csv --log-file $out/launches_cfg2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $out/launches.log 2>&1
# launches matching the regex before the timed resident step: 3 e2e warm-ups x 8 + 3 resident warm-ups x 7 + the work-terms step 7 = 52
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_score|k_prelim_narrow_warp|k_replay|k_setup_queries|k_fold|k_features|k_rows" -s 52 -c 14 -o $out/prof_cfg2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $out/ncu_cfg2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_prelim_wide|k_wide_account" -s 10 -c 2 -o $out/prof_cfg4 python bench.py --workload cfg4 --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $out/ncu_cfg4.log 2>&1
( timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" ) > $out/sanitizer_memcheck.txt 2>&1; echo "memcheck exit $?" >> $out/sanitizer_memcheck.txt
tail -3 $out/sanitizer_memcheck.txt
ls -la $out | head -40
