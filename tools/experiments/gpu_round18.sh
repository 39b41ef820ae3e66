#!/bin/bash
cd "$(dirname "$0")/../.."
tag=${1:-r02_cfold}; out=gpurun_out/$tag; mkdir -p $out
( time timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py ) > $out/tests.log 2>&1
echo "tests exit $?" >> $out/tests.log
tail -4 $out/tests.log
SAGE_B200_LIB=$PWD/sage_b200/lib/ab/phase.so timeout 600 python tools/phase_cycles.py cfg2 2>&1 | tee $out/phase_cfg2.txt
timeout 300 python bench.py --steps 20 --warmup 3 --no-extras > $out/bench_cfg2.json 2> $out/bench_cfg2.err
python - $out/bench_cfg2.json <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1])
e=d["e2e"]
print("value %.3fM e2e %.3fM (%.3f ms, median call %.3f, in-lib %.3f) pageable %.3fM (%.2f ms)" % (d["value"]/1e6, e["value"]/1e6, e["ms_per_step"], e["ms_per_call_median_rank0"], e["ms_in_library_median_rank0"], e["pageable"]["value"]/1e6, e["pageable"]["ms_per_step"]), {k: round(v,3) for k,v in d["phases_ms_per_step"].items()}, d.get("parity_checked"))
PY
