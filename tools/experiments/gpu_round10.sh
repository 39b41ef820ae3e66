#!/bin/bash
cd "$(dirname "$0")/../.."
tag=${1:-r02_k}; out=gpurun_out/$tag; mkdir -p $out
( time timeout 2400 python -m pytest tests -m gpu -q ) > $out/tests.log 2>&1
echo "tests exit $?" >> $out/tests.log
tail -6 $out/tests.log
