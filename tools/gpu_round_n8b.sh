#!/bin/bash
cd "$(dirname "$0")/.."
tag=${1:-r02_n8b}; out=gpurun_out/$tag; mkdir -p $out
( time timeout 600 python tools/bench_multi.py --steps 10 --only 8 --numa-blocks 1 ) > $out/bench_multi_numa.json 2> $out/bench_multi_numa.err
( time timeout 600 python tools/bench_multi.py --steps 10 --only 8 --numa-blocks 0 ) > $out/bench_multi_plain.json 2> $out/bench_multi_plain.err
head -c 1200 $out/bench_multi_numa.json; echo; head -c 1200 $out/bench_multi_plain.json
