#!/bin/bash
cd "$(dirname "$0")/../.."
tag=${1:-r02_final2}; out=gpurun_out/$tag; mkdir -p $out
( time timeout 2400 python -m pytest tests -m gpu -q ) > $out/tests.log 2>&1
echo "tests exit $?" >> $out/tests.log
tail -4 $out/tests.log
( time timeout 900 python bench.py --steps 20 --warmup 3 ) > $out/bench_cfg2_full.json 2> $out/bench_cfg2_full.err
SAGE_B200_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2> $out/trace_cfg2.err
python - $out/bench_cfg2_full.json <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1])
p=d["phases_ms_per_step"]
print("value %.3fM e2e %.3fM (%.2f ms) pageable %.3fM" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"], d["e2e"]["pageable"]["value"]/1e6), {k: round(v,3) for k,v in p.items()}, d["roofline"]["frac"], d["roofline"]["dram_frac"], d["roofline"]["traffic_note"][:60])
for k,v in d.get("extra",{}).items(): print("   extra", k, ("value %.3fM e2e %.3fM" % (v["value"]/1e6, v["e2e"]["value"]/1e6)) if "value" in v else v, v.get("parity_checked",{}).get("psms_identical_to_oracle"))
PY
