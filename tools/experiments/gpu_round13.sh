#!/bin/bash
# e2e timelines of chunked calls (SAGE_B200_TRACE) for several chunkings of the 50k-spectrum batch
cd "$(dirname "$0")/../.."
tag=${1:-r02_chunks}; out=gpurun_out/$tag; mkdir -p $out
for cfg in "SAGE_B200_PIPELINE_CHUNKS=1" "SAGE_B200_PIPELINE_CHUNKS=2" "SAGE_B200_PIPELINE_CHUNKS=3" "SAGE_B200_PIPELINE_CHUNKS=4" "SAGE_B200_FIRST_CHUNK_PCT=20"; do
  n=$(echo $cfg | tr '=' '_')
  env $cfg SAGE_B200_TRACE=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > $out/bench_$n.json 2> $out/trace_$n.err
  python - $out/bench_$n.json $n <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1])
print(sys.argv[2], "e2e %.3fM (%.3f ms) pageable %.3fM" % (d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"], d["e2e"]["pageable"]["value"]/1e6))
PY
  grep "chunk base" $out/trace_$n.err | sed -n '9,14p'
done
