#!/bin/bash
# small-block narrow index, second sweep + one ncu capture of the block kernel
cd "$(dirname "$0")/../.."
tag=${1:-r02_nblk2}; out=gpurun_out/$tag; mkdir -p $out
( time timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py ) > $out/tests.log 2>&1
echo "tests exit $?" >> $out/tests.log
tail -4 $out/tests.log
for cfg in "SAGE_B200_NARROW_BLOCK=512 SAGE_B200_NARROW_CELLS_X=4" "SAGE_B200_NARROW_BLOCK=256 SAGE_B200_NARROW_CELLS_X=4" "SAGE_B200_NARROW_BLOCK=128 SAGE_B200_NARROW_CELLS_X=4" "SAGE_B200_NARROW_BLOCK=256 SAGE_B200_NARROW_CELLS_X=2" "SAGE_B200_NARROW_BLOCK=256 SAGE_B200_NARROW_CELLS_X=8"; do
  n=$(echo $cfg | tr '= ' '__')
  env $cfg timeout 300 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > $out/bench_$n.json 2> $out/bench_$n.err
  python - $out/bench_$n.json "$cfg" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1])
    e=d["e2e"]
    print(sys.argv[2], "value %.3fM e2e %.3fM (%.3f ms, in-lib %.3f)" % (d["value"]/1e6, e["value"]/1e6, e["ms_per_step"], e["ms_in_library_median_rank0"]), {k: round(v,3) for k,v in d["phases_ms_per_step"].items() if k in ("setup","prelim","prelim_count","score")})
except Exception as ex:
    print(sys.argv[2], "failed", ex)
PY
done
SAGE_B200_NARROW_BLOCK=256 SAGE_B200_NARROW_CELLS_X=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_prelim_narrow_warp" -s 8 -c 1 -o $out/prof_nblk python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $out/ncu.log 2>&1
ls -la $out/*.ncu-rep
