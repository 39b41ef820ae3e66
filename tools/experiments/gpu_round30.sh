#!/bin/bash
cd "$(dirname "$0")/../.."
tag=${1:-r02_split4}; out=gpurun_out/$tag; mkdir -p $out
( time timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py ) > $out/tests.log 2>&1
echo "tests exit $?" >> $out/tests.log
tail -3 $out/tests.log | cut -c1-200
for v in main split11; do
  lib=sage_b200/lib/ab/$v.so; [ $v = main ] && lib=sage_b200/lib/libsage_b200.so
  SAGE_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps 20 --warmup 3 --no-extras > $out/bench_$v.json 2> $out/bench_$v.err
  python - $out/bench_$v.json $v <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1])
    e=d["e2e"]
    print(sys.argv[2], "value %.3fM e2e %.3fM (%.3f ms, in-lib %.3f) pageable %.3fM" % (d["value"]/1e6, e["value"]/1e6, e["ms_per_step"], e["ms_in_library_median_rank0"], e["pageable"]["value"]/1e6), {k: round(v,3) for k,v in d["phases_ms_per_step"].items() if k in ("setup","prelim","prelim_count","score")}, d.get("parity_checked",{}).get("psms_identical_to_oracle"))
except Exception as ex:
    print(sys.argv[2], "failed", ex); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 13 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(open("$out/launches.csv")))
hi=next(i for i,r in enumerate(rows) if r and r[0]=="ID"); H=rows[hi]; ki,vi=H.index("Kernel Name"),H.index("Metric Value")
for r in rows[hi+1:hi+14]:
    if len(r)>vi: print(r[ki].split("(")[0][-40:], float(r[vi])/1000.0)
PY
